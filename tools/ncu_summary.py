"""ncu -i X.ncu-rep --page raw --csv  ->  short markdown table of the metrics we reason about."""
import csv, io, subprocess, sys
KEEP = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "lts__t_sectors_srcunit_tex_op_red.sum",
        "lts__t_sectors_srcunit_tex_op_red_lookup_miss.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "lts__t_sectors_srcunit_tex_op_read_lookup_miss.sum",
        "lts__t_sectors_srcunit_ltcfabric.sum", "lts__d_atomic_input_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__grid_size", "launch__block_size", "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
rep, title = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
print(f"# {title}\n")
for r in rows[2:]:
    d = dict(zip(hdr, r))
    print(f"## {d.get('Kernel Name','?')[:110]}\n\n| metric | value | unit |\n|---|---|---|")
    for k in KEEP:
        if k in d:
            print(f"| {k} | {d[k]} | {units[hdr.index(k)]} |")
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    dr = sum(float(d.get(k, "0").replace(",", "")) * scale.get(units[hdr.index(k)], 1.0) for k in ("dram__bytes_read.sum", "dram__bytes_write.sum") if k in d)
    print(f"\nDRAM traffic (read + write): {dr / 1e9:.3f} GB\n")
