"""GPU parity tests proper: the CUDA path through the C ABI against (a) the committed golden
vectors produced by the unmodified reference and (b) the oracle on fresh seeded inputs.
Integer counts must be identical; float64 SSE / beta2Cryptic_weighted are compared bit-for-bit
(0 ulp: IEEE division, no FMA contraction -- tighter than the 1 ulp north_star allows)."""
import pytest

from common import (case_flags, first_diff, gpu_process_rows, load_golden, oracle_process_rows, strip_gene)

pytestmark = pytest.mark.gpu


def test_appendix_a_known_answers(ctx):
    g = load_golden("appendix_a.json.gz")
    for case in g["process"]:
        if "qgene" in case:      # locus filter needs the annotation: covered by the CLI test
            continue
        got = gpu_process_rows(ctx, case)
        d = first_diff(got, strip_gene(case["rows"]))
        assert d is None, "%s: %s" % (case["name"], d)


def test_golden_process_fuzz(ctx):
    cases = load_golden("process_fuzz.json.gz")
    assert len(cases) >= 300
    bad = []
    for case in cases:
        d = first_diff(gpu_process_rows(ctx, case), strip_gene(case["rows"]))
        if d:
            bad.append((case["seed"], d))
    assert not bad, "%d/%d cases differ; first: seed %s %s" % (len(bad), len(cases), bad[0][0], bad[0][1])


def test_golden_process_via_bam(ctx, tmp_path):
    cases = load_golden("process_fuzz.json.gz")[:60]
    for case in cases:
        d = first_diff(gpu_process_rows(ctx, case, via_bam=str(tmp_path)), strip_gene(case["rows"]))
        assert d is None, "seed %s: %s" % (case["seed"], d)


def test_golden_combine_recount(ctx):
    from spliser_b200 import Records
    cases = load_golden("combine_fuzz.json.gz") + load_golden("appendix_a.json.gz")["combine"]
    n = 0
    for case in cases:
        flags = case_flags(case) | 8
        by_sample = {}
        for gap in case["gaps"]:
            by_sample.setdefault(gap["sample"], []).append(gap)
        for s, gaps in by_sample.items():
            reads = [tuple(r) for r in case["samples"][s]["reads"]]
            rec = Records.from_reads(["C"], reads)
            arg = [(0 if g["chrom"] == "C" else -1, g["pos"], g["strand"], g["partners"], g["competitors"]) for g in gaps]
            b1, b2 = ctx.recount_records(rec, 1, arg, flags)
            for g, x, y in zip(gaps, b1, b2):
                n += 1
                assert (int(x), int(y)) == (g["beta1"], g["beta2s"]), (case.get("seed", case.get("name")), g, int(x), int(y))
    assert n > 300


@pytest.mark.parametrize("seed0", [900000, 910000])
def test_live_fuzz_vs_oracle(ctx, seed0):
    from oracle import fuzzgen
    bad = []
    for seed in range(seed0, seed0 + 150):
        case = fuzzgen.gen_case(seed, n_chrom=1 + (seed % 3 == 0), dirty=(seed % 4 == 1), max_reads=60)
        d = first_diff(gpu_process_rows(ctx, case), oracle_process_rows(case))
        if d:
            bad.append((seed, d))
    assert not bad, "%d cases differ; first: seed %s %s" % (len(bad), bad[0][0], bad[0][1])


def test_empty_and_ragged_inputs(ctx):
    import numpy as np
    from spliser_b200 import Junctions, Records
    empty_j = Junctions(np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int32), np.zeros(0, np.int64), np.zeros(0, np.uint8))
    # no junctions, no reads
    t = ctx.process_records(Records.from_reads([], []), 0, empty_j, 0)
    assert len(t) == 0
    # reads but no junctions
    t = ctx.process_records(Records.from_reads(["C"], [("C", 10, 0, "50M")]), 1, empty_j, 0)
    assert len(t) == 0
    # junctions but no reads; reads on a chromosome without junctions; reads without CIGAR
    j = Junctions([0, 0], [100, 100], [200, 300], [5, 3], [ord("+"), ord("+")])
    t = ctx.process_records(Records.from_reads(["C", "D"], [("D", 90, 0, "30M"), ("Z", 5, 0, "10M"), ("C", 95, 0, "*")]), 2, j, 0)
    assert list(t.pos) == [100, 200, 300] and list(t.alpha) == [8, 5, 3] and int(t.beta1.sum()) == 0
    assert [float(x) for x in t.sse] == [1.0, 1.0, 1.0]
