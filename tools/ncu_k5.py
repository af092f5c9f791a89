"""C2 group_by twice (bulk-reduce table update, then the plain 3-RED kernel) for an ncu capture of k_gb_consume:
    ncu --set full --clock-control none --import-source on --kernel-name regex:k_gb_consume -o gpurun_out/r02_k5_bulk python tools/ncu_k5.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import polars_b200 as plb  # noqa: E402
import bench  # noqa: E402

plb.init(0)
key, vi, vf = bench.gen_groupby(100_000_000, 1_000_000, 1)
dkey, dvi, dvf = plb.to_device(key), plb.to_device(vi), plb.to_device(vf)
for bulk in os.environ.get("NCU_K5_MODES", "1,0").split(","):
    os.environ["BL_K5_BULK"] = bulk
    plb.group_by_agg(dkey.view(), [("sum", dvi.view()), ("mean", dvf.view()), ("len", None)], False, location=plb.DEVICE)
plb.sync()
