"""Resident timings of the counting variants on a named workload (development aid; bench.py is the measurement)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spliser_b200  # noqa: E402
from spliser_b200 import synth  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "c2"
reads = int(sys.argv[2]) if len(sys.argv) > 2 else 40_000_000
variants = sys.argv[3].split(",") if len(sys.argv) > 3 else ["fused", "stab"]
cfg = {"c2": synth.config_c2, "c3": synth.config_c3_tile}[name](reads)
t0 = time.time()
w = synth.generate(cfg, cache_dir=os.environ.get("SPLISER_BENCH_CACHE", "/tmp/spliser_bench_cache"))
print("generated %d records in %.1f s" % (len(w.records), time.time() - t0), file=sys.stderr)
out = {}
with spliser_b200.Context(0) as ctx:
    tables = {}
    for v in variants:
        ctx.set_variant(v)
        ctx.resident_load(w.records, len(w.chroms), w.junctions, w.flags)
        ctx.resident_count(3)
        st = ctx.resident_count(10)
        out[v] = {k: st[k] / 10 for k in ("ms_total", "ms_graph_dev", "ms_beta1", "ms_spliced", "ms_final")}
        out[v]["reads_per_s"] = len(w.records) / (st["ms_total"] / 10 * 1e-3)
        t = ctx.resident_fetch()
        tables[v] = (int(t.beta1.sum()), int(t.beta2simple.sum()), int(t.alpha.sum()), float(t.sse.sum()))
    out["checksums"] = tables
print(json.dumps(out))
