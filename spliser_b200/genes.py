"""Annotation loading and gene assignment -- host side, outside the counting path.

Restates createGenes (SpliSER_v0_1_8.py:50-116) without HTSeq (not installable here; the file is parsed natively by
spl_genes_parse) and binary_gene_search (S:118-173) with its quirks (the product runs the native spl_gene_search; the
Python version below is the readable restatement the tests compare it with), because the Gene column of the .SpliSER.tsv is a
pure function of (chromosome, position, strand of the BED row that created the site) and must not
change.  HTSeq.GFF_Reader semantics relied upon: iv.start = GFF start - 1, iv.end = GFF end, name =
value of the first attribute (README.md:64).
"""
from __future__ import annotations

from dataclasses import dataclass


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


class GeneColumns:
    """The genes of an annotation as flat arrays grouped by chromosome (what spl_genes_parse returns):
    genes of chromosome c are rows chrom_off[c] .. chrom_off[c + 1] in list order (S:95)."""

    def __init__(self, chrom_off, left, right, strand_id, strand_texts, names=None, names_blob=None, name_off=None):
        self.chrom_off, self.left, self.right, self.strand_id = chrom_off, left, right, strand_id
        self.strand_texts, self._names = strand_texts, names
        self.names_blob, self.name_off = names_blob, name_off       # UTF-8 blob + offsets as the native parser returns them

    @property
    def names(self):
        """Gene names as a list of str: decoded on first use (the TSV writer takes the blob as it is)."""
        if self._names is None:
            blob, off = self.names_blob, self.name_off.tolist()
            self._names = [blob[off[k]:off[k + 1]].decode() for k in range(len(off) - 1)]
        return self._names

    def name(self, k):
        if self._names is not None:
            return self._names[k]
        return self.names_blob[int(self.name_off[k]):int(self.name_off[k + 1])].decode()


class Annotation:
    """chrom_index: first-appearance order (S:90-92); genes: per chromosome, insort by leftPos (S:95), as Gene objects
    (built on first use when the annotation came from the native parser, which keeps them as GeneColumns)."""

    def __init__(self, chrom_index=None, genes=None, query_gene=None, columns=None):
        self.chrom_index = list(chrom_index) if chrom_index is not None else []
        self._genes = genes if genes is not None or columns is not None else []
        self.query_gene = query_gene
        self.columns = columns

    @property
    def genes(self):
        if self._genes is None:
            c = self.columns
            self._genes = [[Gene(self.chrom_index[ci], c.names[k], int(c.left[k]), int(c.right[k]), c.strand_texts[int(c.strand_id[k])])
                            for k in range(int(c.chrom_off[ci]), int(c.chrom_off[ci + 1]))] for ci in range(len(self.chrom_index))]
        return self._genes

    def genes_of(self, chrom_idx):
        return self.genes[chrom_idx] if chrom_idx < len(self.genes) else []


def load_annotation(path, qgene="All") -> Annotation:
    """createGenes (S:50-116), parsed natively (spl_genes_parse).  The -t/--annotationType argument is ignored by the
    reference (S:82 tests the literal 'gene'), so it is not a parameter here."""
    import ctypes as C

    import numpy as np

    from . import _lib as L
    lib = L.load()
    with open(path, "rb") as fh:
        raw = fh.read()
    if b"\r" in raw:                                   # the reference reads in text mode (universal newlines)
        raw = raw.replace(b"\r\n", b"\n").replace(b"\r", b"\n")
    h = C.c_void_p()
    err = C.create_string_buffer(256)
    rc = lib.spl_genes_parse(raw, len(raw), None if qgene == "All" else str(qgene).encode(), C.byref(h), err, 256)
    if rc != 0:
        raise (OverflowError if rc == -5 else ValueError)(err.value.decode())
    try:
        n, nc = lib.spl_genes_n(h), lib.spl_genes_n_chrom(h)

        def arr(p, cnt, dt):
            return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True) if cnt else np.zeros(0, dt)

        def texts(count, get):
            ln = C.c_int64()
            return [C.string_at(get(h, i, C.byref(ln)), ln.value).decode() for i in range(count)]
        name_off = arr(lib.spl_genes_name_off(h), n + 1, np.int64)
        blob = C.string_at(lib.spl_genes_names(h), int(name_off[-1])) if n else b""
        cols = GeneColumns(arr(lib.spl_genes_chrom_off(h), nc + 1, np.int64), arr(lib.spl_genes_left(h), n, np.int32),
                           arr(lib.spl_genes_right(h), n, np.int32), arr(lib.spl_genes_strand_id(h), n, np.int32),
                           texts(lib.spl_genes_n_strand_texts(h), lib.spl_genes_strand_text), names_blob=blob, name_off=name_off)
        ann = Annotation(texts(nc, lib.spl_genes_chrom_name), None, None, cols)
        k = lib.spl_genes_query(h)
        if k >= 0:
            ci = int(np.searchsorted(cols.chrom_off, k, side="right")) - 1
            ann.query_gene = Gene(ann.chrom_index[ci], cols.name(k), int(cols.left[k]), int(cols.right[k]), cols.strand_texts[int(cols.strand_id[k])])
    finally:
        lib.spl_genes_free(h)
    return ann


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
