"""SpliSER-compatible command line: `process` and `combine` with the reference's flags and byte-identical
.SpliSER.tsv / .combined.tsv output, the counting done by libspliser_b200.so on the GPU.

    python -m spliser_b200.cli process -B x.bam -b x.bed -o out [-A genes.gff] [-c chr] [-g gene -m 500000]
                                       [--isStranded -s rf|fr] [--beta2Cryptic]
    python -m spliser_b200.cli combine -S samples.tsv -o out [-g gene] [--isStranded -s rf|fr] [--beta2Cryptic]

What stays in Python is what the reference does in text: BED12 parsing and its filters (S:255-288), the
annotation and gene lookup (S:50-173), the TSV writers (S:641-664, S:722-740) and the lock-step merge of
`combine` (S:742-917).  The body of `process` (S:710-717) is one spl_process call; every checkBam call of
`combine` (S:903) goes into one batched spl_recount call per sample.
"""
from __future__ import annotations

import argparse
import sys
import timeit
from ast import literal_eval

from . import api
from .bed import parse_bed12
from .genes import NA_NAME, gene_name, load_annotation

VERSION = "v0.1.8 (spliser_b200)"


def _eval_partners(text):
    """The Partners column (`str(dict)` of int -> int, S:662) without ast: `combine` evaluates it three times per row and
    ast.literal_eval was two thirds of its run time.  Anything that is not the plain form goes to literal_eval."""
    t = text.strip()
    if t == "{}":
        return {}
    if t.startswith("{") and t.endswith("}"):
        try:
            out = {}
            for item in t[1:-1].split(","):
                k, v = item.split(":")
                out[int(k)] = int(v)
            return out
        except ValueError:
            pass
    return literal_eval(text)


def _eval_competitors(text):
    """The Competitors column (`str(list)` of int, S:663); same fast path / fallback as _eval_partners."""
    t = text.strip()
    if t == "[]":
        return []
    if t.startswith("[") and t.endswith("]"):
        try:
            return [int(x) for x in t[1:-1].split(",")]
        except ValueError:
            pass
    return literal_eval(text)
PROCESS_HEADER = ("Region\tSite\tStrand\tGene\tSSE\talpha_count\tbeta1_count\tbeta2Simple_count\tbeta2Cryptic_count\t"
                  "beta2Cryptic_weighted\tPartners\tCompetitors\n")
COMBINE_HEADER = ("Sample\tRegion\tSite\tStrand\tGene\tSSE\talpha_count\tbeta1_count\tbeta2Simple_count\tbeta2Cryptic_count\t"
                  "beta2_weighted\tPartners\tCompetitors\n")


# ------------------------------------------------------------------------------------------------ process
def process_table(ctx, bam_path, bed_lines, *, annotation=None, qchrom="All", qgene="All", max_intron=0,
                  is_stranded=False, stranded_type=None, beta2_cryptic=False, records=None):
    """Steps 1-3 of process() (S:710-717).  Returns (chrom_index, SiteTable, strand strings per junction row)."""
    chrom_index = list(annotation.chrom_index) if annotation is not None else []
    bounds = None
    if qgene != "All":
        q = annotation.query_gene if annotation is not None else None
        if q is None:                           # S:283: QUERY_gene is None -> AttributeError in the reference
            raise AttributeError("'NoneType' object has no attribute 'getLeftPos'")
        bounds = (q.left, q.right)
    chroms, junc, sstr = parse_bed12(bed_lines, chrom_index, qchrom, bounds, int(max_intron))
    flags = api.mode_flags(is_stranded, stranded_type, beta2_cryptic)
    if records is not None:
        table = ctx.process_records(records, len(chroms), junc, flags)
    else:
        table = ctx.process_bam(bam_path, chroms, junc, flags)
    return chroms, table, sstr


def write_process_tsv(path, chroms, table, sstr, *, annotation=None, is_stranded=False, beta2_cryptic=False):
    """outputBedFile (S:641-664)."""
    with open(path, "w") as out:
        out.write(PROCESS_HEADER)
        for i in range(len(table)):
            ci, pos = int(table.chrom[i]), int(table.pos[i])
            strand = sstr[int(table.first_line[i])]              # full column-6 text of the row that created the site
            gene = gene_name(annotation, ci, pos, strand, is_stranded) if annotation is not None else NA_NAME
            cols = [chroms[ci], str(pos), strand, gene, "{0:.3f}".format(float(table.sse[i])), str(int(table.alpha[i])),
                    str(int(table.beta1[i])), str(int(table.beta2simple[i]))]
            if beta2_cryptic:
                cols += [str(int(table.beta2cryptic[i])), "{0:.5f}".format(float(table.beta2weighted[i]))]
            else:
                cols += ["NA", "NA"]
            cols += [str(table.partners(i)), str(table.competitors(i))]
            out.write("\t".join(cols) + "\n")


def process(inBAM, inBed, outputPath, qGene="All", qChrom="All", maxIntronSize=0, annotationFile=None, aType="gene",
            isStranded=False, strandedType=None, isbeta2Cryptic=False, ctx=None):
    print("Processing")
    print("Stranded Analysis {}".format(strandedType) if isStranded else "Unstranded Analysis")
    annotation = load_annotation(annotationFile, qGene) if annotationFile is not None else None
    own = ctx is None
    ctx = ctx or api.Context(0)
    try:
        with open(inBed) as fh:
            chroms, table, sstr = process_table(ctx, inBAM, fh, annotation=annotation, qchrom=qChrom, qgene=qGene,
                                                max_intron=maxIntronSize, is_stranded=isStranded, stranded_type=strandedType,
                                                beta2_cryptic=isbeta2Cryptic)
    finally:
        if own:
            ctx.close()
    print("Sites:\t\t\t" + str(len(table)))
    write_process_tsv(outputPath + ".SpliSER.tsv", chroms, table, sstr, annotation=annotation, is_stranded=isStranded,
                      beta2_cryptic=isbeta2Cryptic)


# ------------------------------------------------------------------------------------------------ combine
def _chrom_order(paths):
    """Region order across files: the reference builds a before/after graph and sorts it topologically
    (S:761-789, Graph in Gene_Site_Iter_Graph_v0_1_8.py:358-396)."""
    all_chroms, before_list, after_list = [], [], []
    for p in paths:
        before = "-1"
        with open(p) as fh:
            for idx, line in enumerate(fh):
                if idx == 0:
                    continue
                chrom = line.split("\t")[0]
                if chrom != before:
                    before_list.append(before)
                    after_list.append(chrom)
                    before = chrom
                    if chrom not in all_chroms:
                        all_chroms.insert(0, chrom)
    if not all_chroms:
        print("No genomic regions found - EXITING")
        sys.exit()
    all_chroms.insert(0, "-1")
    graph = {}
    for b, a in zip(before_list, after_list):
        lst = graph.setdefault(b, [])
        if a not in lst:
            lst.append(a)
    visited = [False] * len(all_chroms)
    stack = []

    def visit(region):
        visited[all_chroms.index(region)] = True
        for nxt in graph.get(region, []):
            if not visited[all_chroms.index(nxt)]:
                visit(nxt)
        stack.insert(0, region)

    sys.setrecursionlimit(max(sys.getrecursionlimit(), len(all_chroms) + 1000))
    for i, region in enumerate(all_chroms):
        if not visited[i]:
            visit(region)
    return stack[1:]


class _MergedSite:
    __slots__ = ("chrom", "pos", "gene", "rows", "partners", "competitors", "emit")

    def __init__(self, chrom, pos, gene, n):
        self.chrom, self.pos, self.gene = chrom, pos, gene
        self.rows = [None] * n           # per sample: ("has", vals) or ("gap", gap_id) or None (filtered out)
        self.partners = []               # PartnerCounts keys in insertion order (S:889-892)
        self.competitors = []            # sorted unique (S:894-897)
        self.emit = True


def combine(samplesFile, outputPath, qGene="All", isStranded=False, strandedType="fr", isbeta2Cryptic=False, ctx=None,
            records_by_bam=None):
    """combine (S:742-917).  Pass 1 replays the lock-step merge and collects, per sample, the sites it lacks
    together with the partner / competitor / strand context accumulated from lower-indexed samples only
    (the reference's order dependence); one spl_recount call per sample fills them; pass 2
    writes the rows."""
    print("Combining samples...")
    titles, bed_paths, bam_paths = [], [], []
    with open(samplesFile) as fh:
        for line in fh:
            values = line.split("\t")
            if len(values) == 3:
                titles.append(values[0]); bed_paths.append(values[1]); bam_paths.append(values[2].rstrip())
            else:
                print(str(titles), str(bed_paths), str(bam_paths))
                raise Exception("Samples File contains lines that do not have exactly 3 tab-separated columns")
    n = len(titles)
    chroms_in_order = _chrom_order(bed_paths)
    iters = [open(p) for p in bed_paths]
    for it in iters:
        next(it, None)
    current_vals = [[""] * 11 for _ in range(n)]
    chroms = [""] * n
    iter_go, iter_done = [True] * n, [False] * n
    pos_idx, max_idx = 0, len(chroms_in_order) - 1
    current_chrom = chroms_in_order[pos_idx]
    lowest, lowest_strand = -1, "?"
    merged = []
    gaps = [[] for _ in range(n)]        # per sample: (chrom_name, pos, strand, partners, competitors)
    filled = 0
    while not all(iter_done):
        assoc_gene = ""
        for idx, it in enumerate(iters):
            if not iter_done[idx]:
                if iter_go[idx]:
                    nxt = next(it, None)
                    if nxt is not None:
                        current_vals[idx] = nxt.rstrip().split("\t")
                        chroms[idx] = current_vals[idx][0]
                        iter_go[idx] = False
                    else:
                        iter_done[idx] = True
                        chroms[idx] = None
                if chroms[idx] == current_chrom:
                    pos = int(current_vals[idx][1])
                    strand = current_vals[idx][2]
                    if pos < lowest or lowest == -1 or (isStranded and pos == lowest and strand == "+"):     # S:847
                        lowest, lowest_strand = int(pos), strand
                        assoc_gene = current_vals[idx][3]
        if not all(iter_done):
            if not any(c == current_chrom for c in chroms):                                              # S:857-866
                pos_idx += 1
                current_chrom = chroms_in_order[pos_idx] if pos_idx <= max_idx else None
            else:
                site = _MergedSite(current_chrom, lowest, assoc_gene, n)
                site.emit = (qGene == "All" or qGene == assoc_gene)
                strand_now = ""
                filled_gap = False
                for idx, vals in enumerate(current_vals):
                    if (vals[0] == current_chrom and int(vals[1]) == lowest and not iter_done[idx]
                            and (not isStranded or vals[2] == lowest_strand)):                            # S:870
                        iter_go[idx] = True
                        strand_now = str(vals[2])
                        site.rows[idx] = ("has", vals, strand_now)
                        for key in _eval_partners(str(vals[10])):                                           # S:889-892
                            if key not in site.partners:
                                site.partners.append(key)
                        for c in _eval_competitors(str(vals[11])):                                             # S:894-897
                            if c not in site.competitors:
                                site.competitors.append(c)
                                site.competitors.sort()
                    elif site.emit:                                                                       # S:899-904
                        filled_gap = True
                        site.rows[idx] = ("gap", len(gaps[idx]), strand_now)
                        gaps[idx].append((current_chrom, lowest, strand_now, list(site.partners), list(site.competitors)))
                merged.append(site)
                if filled_gap:
                    filled += 1
            lowest = -1
    for it in iters:
        it.close()

    # ---- one batched re-count per sample
    flags = api.mode_flags(isStranded, strandedType, False, combine=True)
    own = ctx is None
    recount = [None] * n
    if any(gaps):
        ctx = ctx or api.Context(0)
    try:
        for idx in range(n):
            if not gaps[idx]:
                continue
            names = []
            for g in gaps[idx]:
                if g[0] not in names:
                    names.append(g[0])
            arg = [(names.index(g[0]), g[1], g[2], g[3], g[4]) for g in gaps[idx]]
            if records_by_bam is not None:
                recount[idx] = ctx.recount_records(records_by_bam(bam_paths[idx], names), len(names), arg, flags)
            else:
                recount[idx] = ctx.recount_bam(bam_paths[idx], names, arg, flags)
    finally:
        if own and ctx is not None:
            ctx.close()

    # ---- pass 2: rows (outputCombinedLines, S:722-740)
    with open(outputPath + ".combined.tsv", "w") as out:
        out.write(COMBINE_HEADER)
        for site in merged:
            if not site.emit:
                continue
            strand_final = ""
            per = []
            for idx in range(n):
                row = site.rows[idx]
                if row is not None:
                    strand_final = row[2] if row[0] == "has" else strand_final
            # the strand printed for every sample row is the site's final strand (setStrand of the last sample that has it)
            for idx in range(n):
                row = site.rows[idx]
                alpha = beta1 = beta2s = beta2c = 0
                beta2w = 0.0
                pcounts = {}
                if row is not None and row[0] == "has":
                    vals = row[1]
                    alpha, beta1, beta2s = int(vals[5]), int(vals[6]), int(vals[7])
                    if vals[8] != "NA":
                        beta2c, beta2w = int(vals[8]), float(vals[9])
                    pcounts = _eval_partners(str(vals[10]))
                elif row is not None and row[0] == "gap":
                    b1, b2 = recount[idx]
                    beta1, beta2s = int(b1[row[1]]), int(b2[row[1]])
                per.append((alpha, beta1, beta2s, beta2c, beta2w, pcounts, row))
            for idx in range(n):
                alpha, beta1, beta2s, beta2c, beta2w, pcounts, row = per[idx]
                sse = 0.0
                if row is not None and row[0] == "has":                      # calculateSSE, S:626-639 (gap rows: setSSE(0.0))
                    betas = beta1 + beta2s
                    if isbeta2Cryptic:
                        betas = betas + beta2w
                    den = alpha + betas
                    sse = (alpha / den) if den > 0.0 else 0.0
                cols = [titles[idx], site.chrom, str(site.pos), strand_final, site.gene, "{0:.3f}".format(sse), str(alpha),
                        str(beta1), str(beta2s)]
                if isbeta2Cryptic:
                    cols += [str(beta2c), str(beta2w if (row is not None and row[0] == "has" and row[1][8] != "NA") else 0.0)]
                else:
                    cols += ["NA", "NA"]
                cols += [str({k: int(pcounts.get(k, 0)) for k in site.partners}), str(site.competitors)]
                out.write("\t".join(cols) + "\n")
    print("Filled in Beta read counts for {} Sites not detected in some samples".format(filled))


# ------------------------------------------------------------------------------------------------ main
def main(argv=None):
    print("\nSpliSER " + VERSION + "\n")
    start = timeit.default_timer()
    parser = argparse.ArgumentParser(description="SpliSER - Splice Site Strength Estimates from RNA-seq (B200 counting path)")
    sub = parser.add_subparsers(dest="command")
    p = sub.add_parser("process")
    p.add_argument("-B", "--BAMFile", dest="inBAM", required=True)
    p.add_argument("-b", "--bedFile", dest="inBed", required=True)
    p.add_argument("-o", "--outputPath", dest="outputPath", required=True)
    p.add_argument("-A", "--annotationFile", dest="annotationFile", required=False)
    p.add_argument("-t", "--annotationType", dest="aType", nargs="?", default="gene", type=str)
    p.add_argument("-c", "--chromosome", dest="qChrom", nargs="?", default="All", type=str)
    p.add_argument("-g", "--gene", dest="qGene", nargs="?", default="All", type=str)
    p.add_argument("-m", "--maxIntronSize", dest="maxIntronSize", nargs="?", default=0, type=int)
    p.add_argument("--isStranded", dest="isStranded", default=False, action="store_true")
    p.add_argument("-s", "--strandedType", dest="strandedType", nargs="?", type=str)
    p.add_argument("--beta2Cryptic", dest="isbeta2Cryptic", default=False, action="store_true")
    c = sub.add_parser("combine")
    c.add_argument("-S", "--samplesFile", dest="samplesFile", required=True)
    c.add_argument("-o", "--outputPath", dest="outputPath", required=True)
    c.add_argument("-g", "--gene", dest="qGene", nargs="?", default="All", type=str)
    c.add_argument("--isStranded", dest="isStranded", default=False, action="store_true")
    c.add_argument("-s", "--strandedType", dest="strandedType", nargs="?", default="fr", type=str)
    c.add_argument("--beta2Cryptic", dest="isbeta2Cryptic", default=False, action="store_true")
    kwargs = vars(parser.parse_args(argv))
    command = kwargs.pop("command")
    if command == "process" and kwargs.get("qGene") != "All" and (kwargs.get("annotationFile") is None or kwargs.get("maxIntronSize") is None):
        parser.error("--gene requires --annotationFile and --maxIntronSize")                      # S:1350-1353
    elif command in ("process", "combine") and kwargs.get("isStranded") and kwargs.get("strandedType") is None:
        parser.error("--isStranded requires parameter --strandedType/-s as fr or rf")            # S:1354-1355
    elif command == "process":
        process(**kwargs)
    elif command == "combine":
        combine(**kwargs)
    else:
        parser.error("command must be process or combine (combineShallow / output are unchanged Python in the reference)")
    print("Total runtime (s): \t" + str(timeit.default_timer() - start))


if __name__ == "__main__":
    main()
