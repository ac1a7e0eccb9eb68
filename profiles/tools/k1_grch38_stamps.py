"""K1 on a GRCh38-shaped table with few records (the table is what K1 sees): phase timeline of k_gb_coop (SPLISER_K1_STAMPS=1) and the
K1 time of the resident passes.  usage: SPLISER_K1_STAMPS=1 python profiles/tools/k1_grch38_stamps.py [records]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spliser_b200  # noqa: E402
from spliser_b200 import synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8_000_000
w = synth.generate(synth.config_c3_full(n))
with spliser_b200.Context(0) as ctx:
    ctx.resident_load(w.records, len(w.chroms), w.junctions, w.flags)
    ctx.resident_count(3)
    st = ctx.resident_count(10)
    print(json.dumps({"records": len(w.records), "junction_rows": len(w.junctions), "sites": st["n_sites"],
                      "ms_graph_dev": st["ms_graph_dev"] / 10, "ms_total": st["ms_total"] / 10}))
