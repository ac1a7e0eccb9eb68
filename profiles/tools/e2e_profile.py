"""A few spl_process_compact calls on a named workload (development aid: launch list of the e2e call under ncu)."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import spliser_b200  # noqa: E402
from spliser_b200 import api, synth  # noqa: E402

reads = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 2
w = synth.generate(synth.config_c2(reads), cache_dir=os.environ.get("SPLISER_BENCH_CACHE", "/tmp/spliser_bench_cache"))
ck = api.CompactRecords.from_records(w.records, alloc=api.pinned_empty)
with spliser_b200.Context(0) as ctx:
    for _ in range(calls):
        t0 = time.perf_counter()
        t = ctx.process_compact(ck, len(w.chroms), w.junctions, w.flags)
        dt = time.perf_counter() - t0
        st = ctx.stats()
        del t
    print(json.dumps({"ms_call": 1e3 * dt, "n_hot_items": st["n_hot_items"], "n_junc_ops": st["n_junc_ops"], "h2d_bytes": st["h2d_bytes"], "n_parts": st["n_parts"]}))
