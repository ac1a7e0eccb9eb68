"""Python face of the native host text layer (csrc/host_text.cpp): Gene column assignment (binary_gene_search,
SpliSER_v0_1_8.py:118-173), the .SpliSER.tsv writer (outputBedFile, S:641-664) and the `combine` merge driver
(S:742-917).  Strings cross the ABI as string tables (one blob + offsets); string equality the reference relies on
(strand texts, region names, gene names) becomes equality of table ids.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .api import GapTable, SiteTable, SpliserError, _ptr


class StrTable:
    """list[str] -> spl_strtab (keeps the buffers alive)."""

    def __init__(self, strings):
        enc = [s.encode() for s in strings]
        self.blob = b"".join(enc)
        self.off = np.zeros(len(enc) + 1, dtype=np.int64)
        if enc:
            np.cumsum([len(e) for e in enc], out=self.off[1:])
        self.c = L.StrTab(len(enc), self.blob, _ptr(self.off, L.c_i64p))

    @classmethod
    def from_blob(cls, blob, off):
        """UTF-8 blob + int64 offsets (n + 1) as they are, without a round trip through str."""
        t = cls.__new__(cls)
        t.blob = bytes(blob)
        t.off = np.ascontiguousarray(off, dtype=np.int64)
        t.c = L.StrTab(len(t.off) - 1, t.blob, _ptr(t.off, L.c_i64p))
        return t

    def ref(self):
        return C.byref(self.c)


def _columns(t: SiteTable):
    keep = [np.ascontiguousarray(t.chrom, np.int32), np.ascontiguousarray(t.pos, np.int32),
            np.ascontiguousarray(t.first_line, np.int64), np.ascontiguousarray(t.alpha, np.int64),
            np.ascontiguousarray(t.beta1, np.int64), np.ascontiguousarray(t.beta2simple, np.int64),
            np.ascontiguousarray(t.beta2cryptic, np.int64), np.ascontiguousarray(t.beta2weighted, np.float64),
            np.ascontiguousarray(t.sse, np.float64), np.ascontiguousarray(t.partner_off, np.int64),
            np.ascontiguousarray(t.partner_pos, np.int32), np.ascontiguousarray(t.partner_cnt, np.int64),
            np.ascontiguousarray(t.comp_off, np.int64), np.ascontiguousarray(t.comp_pos, np.int32)]
    ty = [L.c_i32p, L.c_i32p, L.c_i64p, L.c_i64p, L.c_i64p, L.c_i64p, L.c_i64p, L.c_f64p, L.c_f64p, L.c_i64p, L.c_i32p,
          L.c_i64p, L.c_i64p, L.c_i32p]
    return keep, L.SiteColumns(len(t), *[_ptr(a, p) for a, p in zip(keep, ty)])


def strand_ids(strand_strings, vocab=None):
    """Distinct texts -> ids in first-appearance order.  Returns (vocab dict, int32 ids)."""
    vocab = {} if vocab is None else vocab
    if not vocab and hasattr(strand_strings, "texts") and hasattr(strand_strings, "ids"):      # bed.StrandColumn
        vocab.update((s, i) for i, s in enumerate(strand_strings.texts))
        if len(vocab) == len(strand_strings.texts):
            return vocab, strand_strings.ids
        vocab.clear()
    ids = np.fromiter((vocab.setdefault(s, len(vocab)) for s in strand_strings), dtype=np.int32, count=len(strand_strings))
    return vocab, ids


class _LazyNames:
    """The flat gene name list of a natively parsed annotation: decoded only if somebody indexes it."""

    def __init__(self, gc):
        self._gc = gc

    def __len__(self):
        return len(self._gc.name_off) - 1

    def __getitem__(self, k):
        return self._gc.names[k]

    def __iter__(self):
        return iter(self._gc.names)


def _gene_columns(annotation, vocab):
    """Per chromosome (left, right, strand id) arrays + the flat name list, built once per annotation and strand vocabulary."""
    key = tuple(sorted(vocab.items()))
    cached = getattr(annotation, "_columns", None)
    if cached is not None and cached[0] == key:
        return cached[1]
    gc = getattr(annotation, "columns", None)
    if gc is not None:                                   # native parser: arrays already there, strand ids -> this vocabulary
        smap = np.array([vocab.setdefault(t, len(vocab)) for t in gc.strand_texts] or [0], dtype=np.int32)
        blob_tab = StrTable.from_blob(gc.names_blob, gc.name_off) if getattr(gc, "names_blob", None) is not None else None
        names, base, cols = (gc if blob_tab is not None else gc.names), gc.chrom_off[:-1].tolist(), []
        for ci in range(len(gc.chrom_off) - 1):
            a, b = int(gc.chrom_off[ci]), int(gc.chrom_off[ci + 1])
            cols.append((np.ascontiguousarray(gc.left[a:b]), np.ascontiguousarray(gc.right[a:b]),
                         np.ascontiguousarray(smap[gc.strand_id[a:b]])))
    else:
        names, base, cols = [], [], []
        for genes in annotation.genes:
            base.append(len(names))
            names.extend(g.name for g in genes)
            cols.append((np.array([g.left for g in genes], dtype=np.int32), np.array([g.right for g in genes], dtype=np.int32),
                         np.array([vocab.setdefault(g.strand, len(vocab)) for g in genes], dtype=np.int32)))
    if gc is not None and blob_tab is not None:
        out = (_LazyNames(gc), base, cols, blob_tab)
    else:
        out = (names, base, cols, StrTable(names))
    try:
        annotation._columns = (tuple(sorted(vocab.items())), out)
    except AttributeError:
        pass
    return out


def assign_genes(annotation, table: SiteTable, site_strand, vocab, is_stranded):
    """Gene column of every site: binary_gene_search over the site's chromosome with the strand text of the BED row that
    created the site (S:313-329).  Returns (gene name list, int32 index per site, -1 = NA)."""
    lib = L.load()
    plus, minus = vocab.setdefault("+", len(vocab)), vocab.setdefault("-", len(vocab))
    names, base, cols, _ = _gene_columns(annotation, vocab)
    out = np.full(len(table), -1, dtype=np.int32)
    chrom = np.asarray(table.chrom)
    pos = np.ascontiguousarray(table.pos, dtype=np.int32)
    site_strand = np.ascontiguousarray(site_strand, dtype=np.int32)
    # sites come grouped by chromosome (output order); fall back to a mask per chromosome when they do not
    change = np.flatnonzero(np.diff(chrom)) + 1 if len(chrom) else np.zeros(0, np.int64)
    starts = np.concatenate(([0], change)).astype(np.int64) if len(chrom) else np.zeros(0, np.int64)
    ends = np.concatenate((change, [len(chrom)])).astype(np.int64) if len(chrom) else np.zeros(0, np.int64)
    for a, b in zip(starts.tolist(), ends.tolist()):
        ci = int(chrom[a])
        if ci < 0 or ci >= len(cols) or not len(cols[ci][0]):
            continue
        gl, gr, gs = cols[ci]
        p, st = pos[a:b], site_strand[a:b]
        idx = np.empty(b - a, dtype=np.int32)
        rc = lib.spl_gene_search(len(gl), _ptr(gl, L.c_i32p), _ptr(gr, L.c_i32p), _ptr(gs, L.c_i32p), b - a,
                                 _ptr(p, L.c_i32p), _ptr(st, L.c_i32p), plus, minus, int(bool(is_stranded)), _ptr(idx, L.c_i32p))
        if rc != 0:
            raise SpliserError("spl_gene_search failed (code %d)" % rc)
        out[a:b] = np.where(idx >= 0, idx + base[ci], -1)
    return names, out


def write_process_tsv(path, chroms, table: SiteTable, strand_strings, *, annotation=None, is_stranded=False,
                      beta2_cryptic=False):
    lib = L.load()
    vocab, line_strand = strand_ids(strand_strings)
    gene_tab, site_gene = None, None
    if annotation is not None and len(table):
        site_strand = line_strand[np.asarray(table.first_line)]
        names, site_gene = assign_genes(annotation, table, site_strand, vocab, is_stranded)
        gene_tab = _gene_columns(annotation, vocab)[3]
    texts = [None] * len(vocab)
    for s, i in vocab.items():
        texts[i] = s
    ctab, stab = StrTable(chroms), StrTable(texts)
    keep, cols = _columns(table)
    err = C.create_string_buffer(512)
    rc = lib.spl_write_process_tsv(str(path).encode(), C.byref(cols), ctab.ref(), stab.ref(), _ptr(line_strand, L.c_i32p),
                                   gene_tab.ref() if gene_tab is not None else None,
                                   _ptr(site_gene, L.c_i32p) if site_gene is not None else None,
                                   int(bool(beta2_cryptic)), err, 512)
    if rc != 0:
        raise (IOError if rc == -3 else SpliserError)("spl_write_process_tsv: %s" % err.value.decode())


class CombineMerge:
    """spl_combine: the merge driver of `combine` around the re-count calls."""

    def __init__(self):
        self._lib = L.load()
        self._h = C.c_void_p()
        if self._lib.spl_combine_create(C.byref(self._h)) != 0:
            raise MemoryError("spl_combine_create")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.spl_combine_destroy(self._h)
            self._h = C.c_void_p()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.spl_combine_last_error(self._h).decode()
            raise (IOError if rc == -3 else SpliserError)("%s: %s (code %d)" % (what, msg, rc))

    def add_sample(self, title, tsv_path):
        self._check(self._lib.spl_combine_add_sample(self._h, str(title).encode(), str(tsv_path).encode()), "spl_combine_add_sample")

    def add_samples(self, titles, tsv_paths, threads=0):
        """All samples of the samples file, in its order; the tables are parsed concurrently."""
        n = len(titles)
        t = (C.c_char_p * max(1, n))(*[str(x).encode() for x in titles])
        p = (C.c_char_p * max(1, n))(*[str(x).encode() for x in tsv_paths])
        self._check(self._lib.spl_combine_add_samples(self._h, n, t, p, int(threads)), "spl_combine_add_samples")

    def set_threads(self, n):
        self._check(self._lib.spl_combine_set_threads(self._h, int(n)), "spl_combine_set_threads")

    def region_names(self):
        return [self._lib.spl_combine_region_name(self._h, i).decode() for i in range(self._lib.spl_combine_n_regions(self._h))]

    def sample_runs(self, k):
        p = L.c_i32p()
        n = self._lib.spl_combine_sample_runs(self._h, k, C.byref(p))
        return [int(p[i]) for i in range(n)]

    def merge(self, region_order, qgene="All", is_stranded=False):
        order = np.ascontiguousarray(region_order, dtype=np.int32)
        q = None if qgene == "All" else str(qgene).encode()
        self._check(self._lib.spl_combine_merge(self._h, len(order), _ptr(order, L.c_i32p), q, int(bool(is_stranded))), "spl_combine_merge")

    def merge_shallow(self, region_order, qgene="All", is_stranded=False, min_samples=0, min_reads=10, min_sse=0.0):
        """combineShallow's merge (S:977-1167): minSamples / minReads / minSSE filters, -g as a row pre-filter."""
        order = np.ascontiguousarray(region_order, dtype=np.int32)
        q = None if qgene == "All" else str(qgene).encode()
        self._check(self._lib.spl_combine_merge_shallow(self._h, len(order), _ptr(order, L.c_i32p), q, int(bool(is_stranded)),
                                                        int(min_samples), int(min_reads), float(min_sse)), "spl_combine_merge_shallow")

    def n_gaps(self, k):
        return int(self._lib.spl_combine_gaps(self._h, k, None, None, None, None, None, None, None))

    def gaps(self, k) -> GapTable:
        """Gap list of sample k; chromosome indices are region ids (region_names())."""
        ps = [L.c_i32p(), L.c_i32p(), L.c_u8p(), L.c_i64p(), L.c_i32p(), L.c_i64p(), L.c_i32p()]
        n = self._lib.spl_combine_gaps(self._h, k, *[C.byref(p) for p in ps])
        if n < 0:
            raise SpliserError("spl_combine_gaps(%d) before merge" % k)

        def arr(p, cnt, dt):
            return np.ctypeslib.as_array(p, shape=(cnt,)).astype(dt, copy=True) if cnt else np.zeros(0, dt)
        p_off = arr(ps[3], n + 1, np.int64)
        c_off = arr(ps[5], n + 1, np.int64)
        return GapTable(arr(ps[0], n, np.int32), arr(ps[1], n, np.int32), arr(ps[2], n, np.uint8),
                        p_off, arr(ps[4], int(p_off[-1]), np.int32), c_off, arr(ps[6], int(c_off[-1]), np.int32))

    def set_recount(self, k, beta1, beta2simple):
        b1 = np.ascontiguousarray(beta1, dtype=np.int64)
        b2 = np.ascontiguousarray(beta2simple, dtype=np.int64)
        self._check(self._lib.spl_combine_set_recount(self._h, k, len(b1), _ptr(b1, L.c_i64p), _ptr(b2, L.c_i64p)), "spl_combine_set_recount")

    def n_filled(self):
        return int(self._lib.spl_combine_n_filled(self._h))

    def n_sites(self):
        return int(self._lib.spl_combine_n_sites(self._h))

    def write(self, path, beta2_cryptic=False):
        self._check(self._lib.spl_combine_write(self._h, str(path).encode(), int(bool(beta2_cryptic))), "spl_combine_write")
