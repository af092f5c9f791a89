"""Writes tests/golden/pdsh_lineitem_head.json from the reference's own sample of the PDS-H `lineitem` table
(/root/reference/examples/datasets/pds_heads/lineitem.feather, 10 rows, the real schema: Int64 / Float64 columns,
large_string flags, timestamp[us] dates).  Run in the authoring container (the reference checkout does not exist on
the GPU box); the JSON travels with the repo.  Only the columns PDS-H Q1 touches are kept."""
import json
import os

import pyarrow.feather as feather

SRC = "/root/reference/examples/datasets/pds_heads/lineitem.feather"
COLS = ["l_quantity", "l_extendedprice", "l_discount", "l_tax", "l_returnflag", "l_linestatus", "l_shipdate"]

t = feather.read_table(SRC).select(COLS)
out = {"source": "examples/datasets/pds_heads/lineitem.feather (pola-rs/polars @ 4db92c1)", "schema": {c: str(t.schema.field(c).type) for c in COLS}, "columns": {}}
for c in COLS:
    col = t.column(c)
    if str(col.type).startswith("timestamp"):
        out["columns"][c] = [int(v.value) if v.is_valid else None for v in col]          # microseconds since epoch
    else:
        out["columns"][c] = col.to_pylist()
with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "pdsh_lineitem_head.json"), "w") as f:
    json.dump(out, f, indent=1)
print("rows", t.num_rows)
