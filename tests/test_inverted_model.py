"""The algorithm the kernels implement (range-add minus exceptions, tests/inverted_model.py) against
the oracle's per-site / per-read restatement -- a CPU check of the design itself."""
import inverted_model as M
from oracle import fuzzgen
from oracle import spliser_oracle as O


def _cls(strand, stranded):
    if not stranded:
        return 0
    return 1 if strand == "+" else 2 if strand == "-" else 3


def test_inverted_model_equals_oracle():
    n = 0
    for seed in range(3000, 3400):
        case = fuzzgen.gen_case(seed, n_chrom=1 + (seed % 3 == 0), dirty=(seed % 4 == 1))
        chroms, junc = O.parse_bed(case["bed"].splitlines(True))
        rbc = [[] for _ in chroms]
        for c, p, f, cg in case["reads"]:
            rbc[chroms.index(c)].append((p, f, O.parse_cigar(cg)))
        st, rf = case["stranded"], case["stype"] == "rf"
        for combine in (False, True):
            sites = O.build_sites(len(chroms), junc, st)
            for ci, arr in enumerate(sites):
                for s in arr:
                    O.check_bam(s, rbc[ci], st, case["stype"], combine)
                g = M.Graph([s.pos for s in arr], [_cls(s.strand, st) for s in arr], [list(s.pcounts) for s in arr],
                            [list(s.comp) for s in arr])
                b1, b2, dc = M.count_chrom(g, rbc[ci], st, rf, combine, want_dc=True)
                assert b1 == [s.beta1 for s in arr], seed
                assert b2 == [s.beta2s for s in arr], seed
                assert dc == {(t, p): v for t, s in enumerate(arr) for p, v in s.dc.items()}, seed
                n += 1
    assert n > 800
