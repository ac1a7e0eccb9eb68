"""Profiling helper: times spl_process from a BAM file (device ingest) on configs[1]; run it under ncu for the per-kernel times."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
import spliser_b200
from spliser_b200 import synth
CACHE = os.path.join(tempfile.gettempdir(), "spliser_bench_cache")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 40_000_000
with_seq = len(sys.argv) > 2 and sys.argv[2] == "seq"            # a sequencer-shaped file (read names, SEQ, QUAL)
w = synth.generate(synth.config_c2(n), cache_dir=CACHE)
bam = os.path.join(CACHE, "prof_%d%s.bam" % (n, "_seq" if with_seq else ""))
if not os.path.exists(bam):
    w.records.write_bam(bam, w.chroms, w.chrom_len, with_seq=with_seq)
ctx = spliser_b200.Context(0)
for i in range(3):
    t = time.perf_counter(); ctx.process_bam(bam, w.chroms, w.junctions, w.flags); dt = time.perf_counter() - t
    st = ctx.stats(); print("iter", i, "ms %.1f" % (1e3 * dt), "ingest %.1f" % st["ms_decode"], "dev", st["bam_on_device"], flush=True)
