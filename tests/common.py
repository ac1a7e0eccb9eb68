"""Shared helpers of the test-suite: golden fixture loading and SiteTable <-> oracle row comparison."""
from __future__ import annotations

import gzip
import json
import os

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    with open(os.path.join(GOLDEN, name), "rb") as fh:
        return json.loads(gzip.decompress(fh.read()))


def case_flags(case):
    f = 0
    if case.get("stranded"):
        f |= 1 | (2 if case.get("stype") == "rf" else 0)
    if case.get("cryptic"):
        f |= 4
    return f


def table_rows(chroms, t):
    """SiteTable -> list of dicts in the layout of oracle.ref_runner._site_dump (minus gene)."""
    rows = []
    for i in range(len(t)):
        rows.append(dict(
            chrom=chroms[int(t.chrom[i])], pos=int(t.pos[i]), strand=t.strand_str(i),
            alpha=int(t.alpha[i]), beta1=int(t.beta1[i]), beta2s=int(t.beta2simple[i]),
            beta2c=int(t.beta2cryptic[i]), beta2w=float(t.beta2weighted[i]).hex(), sse=float(t.sse[i]).hex(),
            partners=[[int(p), int(c)] for p, c in t.partners(i).items()],
            competitors=t.competitors(i)))
    return rows


def strip_gene(rows):
    out = []
    for r in rows:
        r = dict(r)
        r.pop("gene", None)
        r["partners"] = [list(p) for p in r["partners"]]
        out.append(r)
    return out


def first_diff(a, b):
    if len(a) != len(b):
        return "row count %d != %d" % (len(a), len(b))
    for i, (x, y) in enumerate(zip(a, b)):
        if x != y:
            return "row %d:\n  got  %r\n  want %r" % (i, x, y)
    return None


def oracle_process_rows(case):
    """Runs the Python oracle on a fuzz/golden case dict -> rows."""
    from oracle import spliser_oracle as O
    chroms, junc = O.parse_bed(case["bed"].splitlines(True))
    rbc = [[] for _ in chroms]
    for c, p, f, cg in case["reads"]:
        if c in chroms:
            rbc[chroms.index(c)].append((p, f, O.parse_cigar(cg)))
    sites = O.process(len(chroms), junc, rbc, case_flags(case))
    return O.rows_from_sites(chroms, sites)


def gpu_process_rows(ctx, case, via_bam=None):
    """Runs the CUDA path on a case dict -> rows.  via_bam: directory to round-trip the reads through a BAM file."""
    from spliser_b200 import Records
    from spliser_b200.bed import parse_bed12
    chroms, junc, _ = parse_bed12(case["bed"].splitlines(True))
    reads = [tuple(r) for r in case["reads"]]
    if via_bam is None:
        rec = Records.from_reads(chroms, reads)
        t = ctx.process_records(rec, len(chroms), junc, case_flags(case))
    else:
        refs = sorted({r[0] for r in reads} | set(chroms))
        rec = Records.from_reads(refs, reads)
        path = os.path.join(via_bam, "case.bam")
        rec.write_bam(path, refs)
        t = ctx.process_bam(path, chroms, junc, case_flags(case))
    return table_rows(chroms, t)
