"""Genomic-tile sharding of ONE sample, host logic on the CPU: the record slices spliser_b200.dist cuts
for each tile must give, for the sites the tile owns, exactly the counts of the unsharded run -- checked with the oracle,
which (like the reference) counts site by site and so notices any missing alignment."""
import numpy as np
import pytest

from oracle import c_oracle
from spliser_b200 import api, dist, synth


@pytest.mark.parametrize("shape,n_tiles", [("small", 3), ("c3_tile", 4)])
def test_tile_record_slices_reproduce_the_unsharded_counts(shape, n_tiles):
    if shape == "small":
        w = synth.generate(synth.config_small(60_000, seed=101, stranded=True, paired=True))
    else:
        w = synth.generate(synth.config_c3_tile(80_000, tile=2))      # long introns: reads reach far across tile edges
    nc = len(w.chroms)
    full = c_oracle.process(w.records, nc, w.junctions, w.flags | 4, threads=8)
    table = api.build_site_table(nc, w.junctions, w.flags)
    S = len(table)
    assert S == len(full["pos"]) and np.array_equal(table.pos, full["pos"])
    span = dist.max_reference_span(w.records)
    assert span >= 75
    got = {k: np.zeros(S, full[k].dtype) for k in ("beta1", "beta2simple", "beta2cryptic", "sse", "alpha")}
    n_sent = 0
    for t in range(n_tiles):
        rec_t = dist.tile_records(w.records, table, nc, t, n_tiles, max_span=span)
        n_sent += len(rec_t)
        assert len(rec_t) < len(w.records)                              # a tile gets a slice, not the sample
        part = c_oracle.process(rec_t, nc, w.junctions, w.flags | 4, threads=8)
        lo, hi = dist.tile_of(t, n_tiles, S)
        for k in got:
            got[k][lo:hi] = part[k][lo:hi]
    for k in got:
        assert np.array_equal(got[k], full[k]), k
    assert n_sent < 1.5 * len(w.records)                                # edge duplication stays a small fraction


def test_tile_slices_edge_cases():
    from spliser_b200 import Junctions, Records
    rec = Records.from_reads(["A", "B"], [("A", 100, 0, "50M"), ("A", 120, 0, "20M500N20M"), ("B", 10, 0, "30M")])
    j = Junctions([0, 0], [139, 139], [640, 700], [1, 2], [ord("+"), ord("+")])
    table = api.build_site_table(2, j, 0)
    assert dist.max_reference_span(rec) == 540
    # more tiles than sites: the empty tiles get no records at all
    sent = [len(dist.tile_records(rec, table, 2, t, 8)) for t in range(8)]
    assert sum(1 for n in sent if n == 0) >= 5 and max(sent) == 2
    # chromosome B has no site: its read is never sent
    assert all(int(c) == 0 for t in range(8) for c in dist.tile_records(rec, table, 2, t, 8).seg_chrom)
    empty = Records.from_reads(["A"], [])
    assert dist.max_reference_span(empty) == 0 and len(dist.tile_records(empty, table, 2, 0, 2)) == 0


def test_read_balanced_tiles_reproduce_the_unsharded_counts():
    """Tiles cut by read count (dist.balanced_tiles) with per-segment span bounds: the owned slices still concatenate to the
    unsharded result, and no tile receives much more than its share of the records."""
    w = synth.generate(synth.config_c3_tile(80_000, tile=2))
    nc = len(w.chroms)
    full = c_oracle.process(w.records, nc, w.junctions, w.flags | 4, threads=8)
    table = api.build_site_table(nc, w.junctions, w.flags)
    S = len(table)
    n_tiles = 4
    cuts = dist.balanced_tiles(w.records, table, nc, n_tiles)
    assert cuts[0] == 0 and cuts[-1] == S and all(a <= b for a, b in zip(cuts, cuts[1:]))
    spans = dist.segment_max_spans(w.records)
    assert len(spans) == len(w.records.seg_chrom) and max(spans) == dist.max_reference_span(w.records)
    got = {k: np.zeros(S, full[k].dtype) for k in ("beta1", "beta2simple", "beta2cryptic", "sse", "alpha")}
    sent = []
    for t in range(n_tiles):
        lo, hi = cuts[t], cuts[t + 1]
        rec_t = dist.tile_records(w.records, table, nc, t, n_tiles, site_range=(lo, hi), seg_spans=spans)
        sent.append(len(rec_t))
        part = c_oracle.process(rec_t, nc, w.junctions, w.flags | 4, threads=8)
        for k in got:
            got[k][lo:hi] = part[k][lo:hi]
    for k in got:
        assert np.array_equal(got[k], full[k]), k
    assert max(sent) < 1.6 * len(w.records) / n_tiles, sent            # balanced by reads, edge duplicates included


@pytest.mark.parametrize("shape,n_tiles,stranded", [("small", 3, True), ("small", 5, False), ("c3_tile", 4, True)])
def test_tile_junction_rows_give_the_owned_rows_of_the_full_table(shape, n_tiles, stranded):
    """Each tile builds its OWN site table + graph from the junction rows dist.tile_junctions keeps (SURVEY 8(e) has the
    graph replicated; cutting it removes the one per-sample stage that did not shrink with the number of GPUs): the owned
    rows of the per-tile tables, concatenated, equal the unsharded table in every column -- Partners in insertion order,
    PartnerCounts, CompetitorPos, first BED line, beta2Cryptic (which reads the partners' own PartnerCounts) included."""
    if shape == "small":
        w = synth.generate(synth.config_small(60_000, seed=103, stranded=stranded, paired=True))
    else:
        w = synth.generate(synth.config_c3_tile(80_000, tile=2))
    nc = len(w.chroms)
    flags = w.flags | 4
    full = c_oracle.process(w.records, nc, w.junctions, flags, threads=8)
    table = api.build_site_table(nc, w.junctions, w.flags)
    cuts = dist.balanced_tiles(w.records, table, nc, n_tiles)
    spans = dist.segment_max_spans(w.records)
    parts, kept = [], []
    for t in range(n_tiles):
        lo, hi = cuts[t], cuts[t + 1]
        rows, junc_t, (lo2, hi2) = dist.tile_junctions(w.junctions, table, nc, (lo, hi), w.flags)
        kept.append(len(rows))
        rec_t = dist.tile_records(w.records, table, nc, t, n_tiles, site_range=(lo, hi), seg_spans=spans)
        part = c_oracle.process(rec_t, nc, junc_t, flags, threads=8)
        assert hi2 - lo2 == hi - lo and hi2 <= len(part["pos"])
        assert np.array_equal(part["pos"][lo2:hi2], full["pos"][lo:hi]) and np.array_equal(part["strand"][lo2:hi2], full["strand"][lo:hi])

        class T:                                                       # the dict as attributes, like a SiteTable
            pass
        tt = T()
        for k, v in part.items():
            setattr(tt, k, v)
        parts.append(dist.owned_part(tt, (lo2, hi2), rows))
    cat = dist.concat_parts(parts, full)
    assert c_oracle.diff_tables(cat, full) is None
    if n_tiles >= 4 and shape == "c3_tile":
        assert max(kept) < 0.8 * len(w.junctions), kept                # a tile builds a part of the graph, not all of it


def test_tile_junctions_dirty_regime_is_not_cut():
    from spliser_b200 import Junctions
    j = Junctions([0, 0, 0], [100, 100, 900], [300, 300, 1200], [1, 2, 3], [ord("?"), ord("+"), ord("-")])
    table = api.build_site_table(1, j, 1)
    rows, sub, rng = dist.tile_junctions(j, table, 1, (0, 2), 1)
    assert len(rows) == 3 and sub is j and rng == (0, 2)
    # the same rows in an unstranded run are a clean table: the far row is dropped for a tile that owns the first two sites
    table = api.build_site_table(1, j, 0)
    rows, sub, rng = dist.tile_junctions(j, table, 1, (0, 2), 0)
    assert list(rows) == [0, 1] and rng == (0, 2)
