"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.json.gz by running the UNMODIFIED
reference (oracle/ref_runner.py) in the authoring container.  The reference cannot travel to
the GPU box, so its outputs are committed as fixtures together with this script.

    python -m oracle.make_golden            # rewrites tests/golden/

Fixtures:
  appendix_a.json.gz   SURVEY.md Appendix A scenarios A.1-A.4 (known-answer, incl. TSV text)
  process_fuzz.json.gz seeded fuzz cases (oracle/fuzzgen.py) with the reference's per-site rows
  combine_fuzz.json.gz seeded combine cases: per-sample TSVs, combined TSV, every gap re-count
  combine_wide.json.gz same over several regions, with annotation / -g / --beta2Cryptic (fuzzgen.gen_combine_wide_case)
"""
from __future__ import annotations

import gzip
import json
import os
import sys

from . import fuzzgen, ref_runner as R

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

N_PROCESS = 360
N_COMBINE = 80
N_COMBINE_WIDE = 60


def _dump(name, obj):
    os.makedirs(OUT, exist_ok=True)
    raw = json.dumps(obj, separators=(",", ":"), sort_keys=True).encode()
    with open(os.path.join(OUT, name), "wb") as fh:
        fh.write(gzip.compress(raw, mtime=0))
    print(name, len(raw), "bytes raw")


def appendix_a():
    out = []
    C = "C"
    bl = R.bed_line
    bed1 = bl(C, 100, 200, 5, "?") + bl(C, 100, 300, 3, "?") + bl(C, 400, 500, 2, "?")
    reads1 = ([(C, 81, 0, "20M100N20M")] * 5 + [(C, 81, 0, "20M200N20M")] * 3 +
              [(C, 90, 0, "30M"), (C, 190, 0, "30M"), (C, 82, 0, "20M"), (C, 81, 0, "20M"),
               (C, 101, 0, "20M"), (C, 90, 0, "11M1I20M"), (C, 190, 0, "10M2D20M"),
               (C, 131, 0, "20M100N20M")] + [(C, 381, 0, "20M100N20M")] * 2)
    tsv, rows = R.run_process(bed1, reads1)
    out.append(dict(name="A.1", bed=bed1, reads=reads1, stranded=False, stype=None, cryptic=False,
                    tsv=tsv, rows=rows))
    bed2 = bl(C, 100, 300, 4, "+") + bl(C, 250, 300, 2, "+") + bl(C, 100, 300, 1, "-")
    reads2 = ([(C, 81, 0, "20M200N20M")] * 4 + [(C, 231, 0, "20M50N20M")] * 2 +
              [(C, 81, 16, "20M200N20M"), (C, 90, 0, "161M50N20M"), (C, 90, 16, "161M50N20M")] +
              [(C, 90, f, "30M") for f in (0, 16, 99, 147, 163, 83, 1)] + [(C, 240, 0, "30M")])
    for tag, kw in (("unstranded", {}), ("fr", dict(stranded=True, stype="fr")),
                    ("rf", dict(stranded=True, stype="rf"))):
        tsv, rows = R.run_process(bed2, reads2, cryptic=True, **kw)
        out.append(dict(name="A.2-" + tag, bed=bed2, reads=reads2, stranded=kw.get("stranded", False),
                        stype=kw.get("stype"), cryptic=True, tsv=tsv, rows=rows))
    # A.4 locus filter + gene assignment
    gff = ("C\tx\tgene\t90\t320\t.\t+\t.\tID=G1;Name=g1\n"
           "C\tx\tmRNA\t90\t320\t.\t+\t.\tID=G1.1;Parent=G1\n"
           "C\tx\tgene\t380\t520\t.\t-\t.\tID=G2\n")
    tsv, rows = R.run_process(bed1, reads1, gff_text=gff)
    out.append(dict(name="A.4-annot", bed=bed1, reads=reads1, stranded=False, stype=None, cryptic=False,
                    gff=gff, tsv=tsv, rows=rows))
    tsv, rows = R.run_process(bed1, reads1, gff_text=gff, qchrom="C", qgene="G1", max_intron=50, cryptic=True)
    out.append(dict(name="A.4-locus", bed=bed1, reads=reads1, stranded=False, stype=None, cryptic=True,
                    gff=gff, qchrom="C", qgene="G1", max_intron=50, tsv=tsv, rows=rows))
    # dirty regime example from SURVEY.md 8(a): '?', '+', '-', '?' lines at (100,300)
    bed5 = bl(C, 100, 300, 1, "?") + bl(C, 100, 300, 2, "+") + bl(C, 100, 300, 4, "-") + bl(C, 100, 300, 8, "?")
    tsv, rows = R.run_process(bed5, reads2, stranded=True, stype="fr", cryptic=True)
    out.append(dict(name="dirty-1", bed=bed5, reads=reads2, stranded=True, stype="fr", cryptic=True,
                    tsv=tsv, rows=rows))
    # A.3 combine order dependence
    bedA = bl(C, 100, 300, 4, "+") + bl(C, 250, 300, 2, "+")
    readsA = [(C, 81, 0, "20M200N20M")] * 4 + [(C, 231, 0, "20M50N20M")] * 2
    bedB = bl(C, 600, 700, 1, "+")
    readsB = [(C, 81, 0, "20M200N20M"), (C, 90, 0, "161M50N20M"), (C, 90, 0, "30M"), (C, 90, 16, "30M"),
              (C, 581, 0, "20M100N20M")]
    comb = []
    for stranded in (False, True):
        kw = dict(stranded=stranded, stype="fr" if stranded else None)
        tA, _ = R.run_process(bedA, readsA, **kw)
        tB, _ = R.run_process(bedB, readsB, **kw)
        for order in ("AB", "BA"):
            samples = [("A", tA, readsA), ("B", tB, readsB)]
            if order == "BA":
                samples.reverse()
            ctsv, gaps = R.run_combine(samples, stranded=stranded, stype="fr")
            comb.append(dict(name="A.3-%s-%s" % (order, "fr" if stranded else "un"), stranded=stranded, stype="fr",
                             samples=[dict(title=t, tsv=x, reads=r) for t, x, r in samples],
                             combined=ctsv, gaps=gaps))
    return dict(process=out, combine=comb)


def process_fuzz(n=N_PROCESS):
    cases = []
    for seed in range(n):
        case = fuzzgen.gen_case(20260000 + seed, n_chrom=1 + (seed % 3 == 0) + (seed % 11 == 0),
                                dirty=(seed % 4 == 1))
        tsv, rows = R.run_process(case["bed"], case["reads"], stranded=case["stranded"],
                                  stype=case["stype"], cryptic=case["cryptic"])
        case["rows"] = rows
        case["tsv"] = tsv
        cases.append(case)
    return cases


def combine_fuzz(n=N_COMBINE):
    cases = []
    for seed in range(n):
        case = fuzzgen.gen_combine_case(20261000 + seed)
        samples = []
        for s in case["samples"]:
            tsv, _ = R.run_process(s["bed"], s["reads"], stranded=case["stranded"],
                                   stype=case["stype"] if case["stranded"] else None)
            s["tsv"] = tsv
            samples.append((s["title"], tsv, s["reads"]))
        ctsv, gaps = R.run_combine(samples, stranded=case["stranded"], stype=case["stype"])
        case["combined"] = ctsv
        case["gaps"] = gaps
        cases.append(case)
    return cases


def combine_wide(n=N_COMBINE_WIDE):
    cases = []
    for seed in range(n):
        case = fuzzgen.gen_combine_wide_case(20262000 + seed)
        samples = []
        for s in case["samples"]:
            tsv, _ = R.run_process(s["bed"], s["reads"], stranded=case["stranded"],
                                   stype=case["stype"] if case["stranded"] else None, cryptic=case["cryptic"],
                                   gff_text=case["gff"])
            s["tsv"] = tsv
            samples.append((s["title"], tsv, s["reads"]))
        ctsv, gaps = R.run_combine(samples, stranded=case["stranded"], stype=case["stype"], cryptic=case["cryptic"],
                                   qgene=case["qgene"])
        case["combined"] = ctsv
        case["gaps"] = gaps
        cases.append(case)
    return cases


SHALLOW_SETTINGS = ((0, 10, 0.0), (1, 1, 0.0), (2, 2, 0.0), (2, 1, 0.5), (3, 4, 0.25), (1, 3, 0.9))


def combine_shallow():
    """`combineShallow` (SpliSER_v0_1_8.py:920-1167) of the unmodified reference on the samples of the combine_wide and
    combine_fuzz cases (looked up by file + index, not stored again), over a grid of -m / -r / -e settings, with and without -g."""
    import gzip
    import json

    def load(name):
        with open(os.path.join(OUT, name), "rb") as fh:
            return json.loads(gzip.decompress(fh.read()))
    out = []
    for base, step in (("combine_wide.json.gz", 1), ("combine_fuzz.json.gz", 2)):
        cases = load(base)
        for i in range(0, len(cases), step):
            case = cases[i]
            samples = [(s["title"], s["tsv"], [tuple(r) for r in s["reads"]]) for s in case["samples"]]
            genes = sorted({ln.split("\t")[3] for s in case["samples"] for ln in s["tsv"].splitlines()[1:]} - {"NA", ""})
            for j, (ms, mr, me) in enumerate(SHALLOW_SETTINGS):
                if (i + j) % 3 == 2:
                    continue
                if ms == 3:
                    ms = len(samples)
                qgene = genes[(i + j) % len(genes)] if genes and (i + j) % 2 == 1 else "All"
                cryptic = bool(case.get("cryptic", False)) and j % 2 == 0
                ctsv, gaps = R.run_combine(samples, stranded=case["stranded"], stype=case["stype"], cryptic=cryptic, qgene=qgene,
                                           shallow=(ms, mr, me))
                out.append(dict(base=base, index=i, min_samples=ms, min_reads=mr, min_sse=me, qgene=qgene, cryptic=cryptic,
                                combined=ctsv, n_gaps=len(gaps)))
    return out


def combine_shallow_crafted():
    """Known-answer inputs for the quirks of combineShallow's loop that random cases rarely reach:
    X  a dropped position moves on every sample whose current row has that position NUMBER, also a row of another region
       (S:1158-1160): sample B loses its K2:100 / K2:300 rows while K1:100 / K1:300 of sample A are dropped;
    Y  the minSamples count runs over the rows of both strands of a position (S:1079-1084): the weak '+' site of sample A
       is kept because sample B's '-' row at the same position passes;
    Z  a '+' row takes a tied position over from a '-' row and restarts the count (S:1066-1077): sample B's strong '-' row
       is dropped together with sample A's weak '+' row."""
    bl = R.bed_line
    j1, j2 = "20M200N20M", "20M200N20M"
    out = []

    def case(name, chroms, specs, stranded, settings):
        stype = "fr"
        samples = []
        for title, bed, reads in specs:
            tsv, _ = R.run_process(bed, reads, stranded=stranded, stype=stype if stranded else None)
            if not bed:                                   # a sample without junctions: its table is the header line alone
                assert tsv.count("\n") == 1
            samples.append(dict(title=title, tsv=tsv, reads=reads))
        for ms, mr, me in settings:
            ctsv, gaps = R.run_combine([(x["title"], x["tsv"], x["reads"]) for x in samples], stranded=stranded, stype=stype,
                                       shallow=(ms, mr, me))
            out.append(dict(base=None, name=name, chroms=chroms, samples=samples, stranded=stranded, stype=stype, min_samples=ms,
                            min_reads=mr, min_sse=me, qgene="All", cryptic=False, combined=ctsv, n_gaps=len(gaps)))
    a = ("A", bl("K1", 100, 300, 1, "+") + bl("K2", 500, 700, 9, "+"), [("K1", 81, 0, j1)] + [("K2", 481, 0, j2)] * 9)
    b = ("B", bl("K2", 100, 300, 9, "+") + bl("K2", 500, 700, 9, "+"), [("K2", 81, 0, j1)] * 9 + [("K2", 481, 0, j2)] * 9)
    case("X-cross-region-skip", ["K1", "K2"], [a, b], False, [(1, 5, 0.0), (1, 1, 0.0), (2, 5, 0.0), (0, 5, 0.0)])
    case("X-cross-region-skip-BA", ["K1", "K2"], [b, a], False, [(1, 5, 0.0), (2, 1, 0.0)])
    ap = ("A", bl("C", 100, 300, 1, "+") + bl("C", 500, 700, 9, "+"), [("C", 81, 0, j1)] + [("C", 481, 0, j2)] * 9)
    bm = ("B", bl("C", 100, 300, 9, "-") + bl("C", 500, 700, 9, "+"), [("C", 81, 16, j1)] * 9 + [("C", 481, 0, j2)] * 9)
    case("Y-both-strands-counted", ["C"], [ap, bm], True, [(1, 5, 0.0), (2, 5, 0.0), (1, 1, 0.0)])
    case("Z-plus-takes-the-tie", ["C"], [bm, ap], True, [(1, 5, 0.0), (2, 5, 0.0), (1, 1, 0.0), (1, 5, 0.5)])
    case("YZ-unstranded", ["C"], [bm, ap], False, [(1, 5, 0.0), (2, 5, 0.0)])
    # E  a sample whose table holds the header line only (no junction reached the BED12): every site of the others is a gap in it
    empty = ("E", "", [("C", 81, 0, j1)] * 3 + [("C", 90, 0, "30M")])
    case("E-empty-table-last", ["C"], [ap, empty], False, [(0, 10, 0.0), (1, 3, 0.0), (2, 3, 0.0)])
    case("E-empty-table-first", ["C"], [empty, ap], True, [(0, 10, 0.0), (1, 3, 0.0)])
    return out


def main():
    if "--shallow" in sys.argv:
        if not R.reference_available():
            sys.exit("reference not mounted; golden vectors can only be regenerated in the authoring container")
        return _dump("combine_shallow.json.gz", combine_shallow_crafted() + combine_shallow())
    if not R.reference_available():
        sys.exit("reference not mounted; golden vectors can only be regenerated in the authoring container")
    _dump("appendix_a.json.gz", appendix_a())
    _dump("process_fuzz.json.gz", process_fuzz())
    _dump("combine_fuzz.json.gz", combine_fuzz())
    _dump("combine_wide.json.gz", combine_wide())


if __name__ == "__main__":
    main()
