"""Annotation loading and gene assignment -- host-side Python, outside the accelerated path.

Restates createGenes (SpliSER_v0_1_8.py:50-116) without HTSeq (not installable here) and
binary_gene_search (S:118-173) with its quirks, because the Gene column of the .SpliSER.tsv is a
pure function of (chromosome, position, strand of the BED row that created the site) and must not
change.  HTSeq.GFF_Reader semantics relied upon: iv.start = GFF start - 1, iv.end = GFF end, name =
value of the first attribute (README.md:64).
"""
from __future__ import annotations

from dataclasses import dataclass, field


@dataclass(order=False)
class Gene:
    chrom: str
    name: str
    left: int
    right: int
    strand: str

    def __lt__(self, other):                  # Gene.__lt__: ordered by leftPos only (G:18-19)
        return self.left < other.left


NA_NAME = "NA"


@dataclass
class Annotation:
    chrom_index: list = field(default_factory=list)       # first-appearance order (S:90-92)
    genes: list = field(default_factory=list)             # per chromosome, insort by leftPos (S:95)
    query_gene: Gene | None = None

    def genes_of(self, chrom_idx):
        return self.genes[chrom_idx] if chrom_idx < len(self.genes) else []


def _first_attribute(col9: str) -> str:
    first = col9.split(";")[0].strip()
    if "=" in first:
        return first.split("=", 1)[1]
    parts = first.split(None, 1)             # GTF: key "value"
    return parts[1].strip('"') if len(parts) > 1 else first


def load_annotation(path, qgene="All") -> Annotation:
    """createGenes (S:50-116).  The -t/--annotationType argument is ignored by the reference (S:82 tests
    the literal 'gene'), so it is not a parameter here."""
    ann = Annotation()
    index = {}
    with open(path) as fh:
        for line in fh:
            if line.startswith("#") or "\tgene\t" not in line:      # cheap pre-filter; the column test below decides
                continue
            f = line.rstrip("\n").split("\t")
            if len(f) < 9 or f[2] != "gene":
                continue
            chrom = f[0]
            name = _first_attribute(f[8])
            ci = index.get(chrom)
            if ci is None:
                ci = index[chrom] = len(ann.chrom_index)
                ann.chrom_index.append(chrom)
                ann.genes.append([])
            g = Gene(chrom, name, int(f[3]) - 1, int(f[4]), f[6])
            if qgene == "All":
                ann.genes[ci].append(g)
            elif name == qgene:
                ann.query_gene = g
                ann.genes[ci].append(g)
    if qgene == "All":
        for genes in ann.genes:          # insort_right by leftPos in file order (S:95) == stable sort by leftPos
            genes.sort(key=_left_of)
    return ann


def _left_of(g):
    return g.left


def binary_gene_search(array, pos, strand, is_stranded) -> int:
    """S:118-173, control flow kept (overlapping genes make the bisection order-dependent; the final
    +-3 window skips the last gene of the list because of `idx+i < len(array)-1`)."""
    length = len(array)
    if length == 0:
        return -1
    idx = length // 2
    past_max, past_min, last_idx, new_idx = length, 0, -1, idx
    stuck = found = False

    def strand_ok(g):
        return strand == g.strand or (not is_stranded) or (strand != "+" and strand != "-")

    while not stuck and not found:
        g = array[idx]
        if g.left <= pos <= g.right and strand_ok(g):
            found = True
            break
        elif pos >= g.right:
            new_idx = idx + ((past_max - idx) // 2)
            past_min = idx
        elif pos <= g.left:
            new_idx = idx - ((idx - past_min) // 2)
            past_max = idx
            if idx == 1:
                new_idx = 0
        if idx != last_idx:
            last_idx = idx
            idx = new_idx
        else:
            stuck = True
    if not found and stuck:
        for i in range(-3, 3):
            k = idx + i
            if 0 <= k < length - 1 and array[k].left <= pos <= array[k].right:
                if strand_ok(array[k]):
                    found, stuck, idx = True, False, k
    return idx if (found and not stuck) else -1


def gene_name(ann: Annotation | None, chrom_idx: int, pos: int, strand: str, is_stranded: bool) -> str:
    if ann is None:
        return NA_NAME
    arr = ann.genes_of(chrom_idx)
    k = binary_gene_search(arr, pos, strand, is_stranded)
    return arr[k].name if k >= 0 else NA_NAME
