"""bench.py contract (task statement, "Measurement"): the reference arm runs on CPU only — it is the one bench leg
that can be exercised here.  Checks the JSON line the driver parses: one line, the contract's keys, the
`impl: reference` / `cpu_baseline` / `e2e` shape, and that the C2 / C3 variants flags are accepted."""
import json
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline", "dtype", "data", "config",
        "cpu_baseline", "e2e"}


def run_bench(*args):
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1", "--rows", "200000", "--keys", "1000", "--build-rows", "20000", *args],
                       capture_output=True, text=True, timeout=300, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1, f"expected ONE JSON line on stdout, got {len(lines)}"
    return json.loads(lines[0])


@pytest.mark.parametrize("args,metric", [((), "group_by_agg_rows_per_sec"), (("--workload", "groupby", "--skew", "zipf"), "group_by_agg_rows_per_sec"),
                                         (("--workload", "join", "--hit-frac", "0.5", "--dup", "4"), "hash_join_probe_rows_per_sec")])
def test_reference_arm_line(args, metric):
    d = run_bench(*args)
    assert KEYS <= set(d), sorted(KEYS - set(d))
    assert d["impl"] == "reference" and d["metric"] == metric and d["unit"] == "rows/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["ms_per_step"] > 0 and d["vs_baseline"] is None
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "rows/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["config"]["same_config"] is True and d["config"]["rows_per_step"] == 200000
    if not args:      # the default run covers both halves of BASELINE.json's metric: group_by primary, join secondary
        assert [x["metric"] for x in d["secondary"]] == ["hash_join_probe_rows_per_sec"] and d["secondary"][0]["value"] > 0


def test_reference_arm_other_ranks_stay_silent():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1", "--warmup", "1", "--rows", "100000", "--keys", "1000", "--build-rows", "10000"],
                       capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""
