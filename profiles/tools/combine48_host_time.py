"""Host stages of `combine` / `combineShallow` on the 48-sample configs[3]-shaped workload (oracle/c4_shape.py c4x48), no GPU:
the product's native merge driver (parse 48 tables, lock-step merge + gap lists, write) is timed with the re-count calls taken
out (a stand-in context answers them from the C oracle; their time is subtracted and reported apart).  The reference's own
figure for the same files is in tests/golden/c4_48_samples_reference.json (`reference_seconds`).

    python profiles/tools/combine48_host_time.py > profiles/r1d_combine48_host.json
"""
import json
import os
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    import contextlib
    import io
    from oracle import c4_shape
    from spliser_b200 import cli
    from test_cli_cpu import OracleContext
    shape = c4_shape.FORTY_EIGHT
    gold = json.load(open(shape.golden))

    class Timed(OracleContext):
        t = 0.0
        calls = 0

        def recount_bam(self, *a):
            t0 = time.perf_counter()
            try:
                return super().recount_bam(*a)
            finally:
                Timed.t += time.perf_counter() - t0
                Timed.calls += 1

    with tempfile.TemporaryDirectory() as td, contextlib.redirect_stdout(io.StringIO()):
        c4_shape.run_cli(cli, OracleContext(), td, shape)              # writes the 48 tables + samples.tsv (and warms up)
        sf = os.path.join(td, "samples.tsv")
        runs = {"combine": [], "combineShallow": []}
        for _ in range(5):
            for name in runs:
                Timed.t, Timed.calls = 0.0, 0
                t0 = time.perf_counter()
                if name == "combine":
                    cli.combine(sf, os.path.join(td, "t"), isStranded=True, strandedType="rf", ctx=Timed())
                else:
                    cli.combineShallow(sf, os.path.join(td, "t"), isStranded=True, minSamples=shape.shallow[0], minReads=shape.shallow[1],
                                       minSSE=shape.shallow[2], strandedType="rf", ctx=Timed())
                total = time.perf_counter() - t0
                runs[name].append((total - Timed.t, Timed.t, Timed.calls))
    out = {"workload": gold["workload"], "threads": os.cpu_count(), "where": "authoring container (no GPU)",
           "reference_seconds": {"combine": gold["reference_seconds"]["combine"], "combineShallow": gold["reference_seconds"]["combineShallow"],
                                 "note": "unmodified reference, merge + its per-gap checkBam against an in-process read store"}}
    for name, r in runs.items():
        best = min(r)
        out[name] = {"merge_driver_ms_best_of_5": round(best[0] * 1e3, 1), "merge_driver_ms_all": [round(x[0] * 1e3, 1) for x in r],
                     "recount_calls": best[2], "stand_in_recount_ms": round(best[1] * 1e3, 1),
                     "rows": gold["combined_rows" if name == "combine" else "shallow_rows"],
                     "gaps": gold["recounted_gaps" if name == "combine" else "shallow_recounted_gaps"]}
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
