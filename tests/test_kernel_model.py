"""The step-for-step model of the CUDA tile algorithms (tests/kernel_model.py) against the oracle."""
import numpy as np
import pytest

import datagen
import kernel_model as km
import oracle_lib as orc

TW = 7936


def _cases():
    yield "zeros", np.zeros(2 * TW + 100, dtype=np.uint32)
    yield "ones", np.full(TW + 992, 0xFFFFFFFF, dtype=np.uint32)
    yield "dense", datagen.uniform(TW + 500, 0.5, 1)
    yield "sparse", datagen.uniform(3 * TW + 17, 0.001, 2)
    yield "d16", datagen.uniform(2 * TW, 1 / 16, 3)
    yield "clustered", datagen.clustered(3 * TW + 1, 0.3, 300, 4)
    yield "clustered_long", datagen.clustered(4 * TW, 0.01, 2000, 5)
    yield "mix", datagen.group_mix(2 * TW + 31, 0.4, 0.3, 6)
    yield "mix_runs", datagen.group_mix(3 * TW, 0.45, 0.45, 7, run=40)
    yield "alternating_fills", datagen.group_mix(TW + 62, 0.5, 0.5, 8)
    for n in (1, 2, 30, 31, 32, 33, 991, 992, 993, TW - 1, TW, TW + 1):
        yield f"tail_{n}", datagen.uniform(n, 0.02, 100 + n)
        yield f"tailz_{n}", np.zeros(n, dtype=np.uint32)


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name,data", list(_cases()), ids=[c[0] for c in _cases()])
def test_compress_model_matches_oracle(name, data, mode):
    want = orc.compress(data, mode)
    got, _ = km.compress_model([data.tolist()], mode, seed=hash(name) & 0xFFFF)
    assert len(got) == want.size
    assert np.array_equal(np.array(got, dtype=np.uint32), want)


@pytest.mark.parametrize("mode", [0, 1])
def test_compress_model_batch(mode):
    cols = np.stack([
        datagen.uniform(TW + 40, 0.001, 11), np.zeros(TW + 40, dtype=np.uint32),
        datagen.clustered(TW + 40, 0.2, 500, 12), np.zeros(TW + 40, dtype=np.uint32),
        np.full(TW + 40, 0xFFFFFFFF, dtype=np.uint32),
    ])
    want, offs = orc.compress_batch(cols, mode)
    got, got_offs = km.compress_model([c.tolist() for c in cols], mode, seed=5)
    assert np.array_equal(np.array(got, dtype=np.uint32), want)
    assert got_offs == offs.tolist()


@pytest.mark.parametrize("block_mode", [True, False])
@pytest.mark.parametrize("grid,n_tiles,tiles_per_col,p_tail", [
    (1, 40, 40, 0.5), (2, 9, 9, 0.0), (3, 100, 25, 0.3), (7, 1000, 1000, 0.02), (296, 2500, 2500, 0.01),
    (296, 1200, 37, 0.1), (321, 1500, 1500, 0.001), (700, 3000, 500, 0.005), (1000, 999, 999, 0.0),
])
def test_chained_offsets_model_is_a_prefix_sum_with_run_carry(grid, n_tiles, tiles_per_col, p_tail, block_mode):
    """The control warps' chained sum (windows of 320 descriptors, nearest first, two warps handing the CTA's running
    totals to each other) against the plain left-to-right definition of what it computes."""
    rng = np.random.default_rng(grid * 7919 + n_tiles)
    aggs = []
    for k in range(n_tiles):
        has = int(rng.random() < p_tail)
        cnt = int(rng.integers(0, 8193)) if has else 0
        opn = int(rng.integers(0, 8192)) if has else 8192
        if block_mode or k % tiles_per_col == tiles_per_col - 1:
            has, opn = 1, 0
        aggs.append((cnt, opn, has))
    got = km.chained_offsets(aggs, grid, tiles_per_col, block_mode)
    excl = run = 0
    for k, (cnt, opn, has) in enumerate(aggs):
        if block_mode or k % tiles_per_col == 0:
            run = 0
        assert got[k] == (excl, run), (k, got[k], excl, run)
        excl += cnt
        run = opn if has else run + opn


def test_compress_model_is_the_same_for_every_grid():
    data = datagen.clustered(6 * TW + 123, 0.05, 3000, 21)
    want = orc.compress(data, 1)
    for grid in (1, 2, 5, 296):
        got, _ = km.compress_model([data.tolist()], 1, grid=grid)
        assert np.array_equal(np.array(got, dtype=np.uint32), want), grid


def test_compress_model_append_merges_seam():
    # CANONICAL stream compressed as two launches: the second one's leading run joins the last word (launch_seam)
    a = np.zeros(992 * 3, dtype=np.uint32)
    b = np.zeros(992 * 2, dtype=np.uint32)
    b[-1] = 5
    first, _ = km.compress_model([a.tolist()], 1)
    both, _ = km.compress_model([b.tolist()], 1, merge_prev_words=first)
    want = orc.compress(np.concatenate([a, b]), 1)
    assert np.array_equal(np.array(both, dtype=np.uint32), want)
    # no merge when types differ
    a2 = np.full(992, 0xFFFFFFFF, dtype=np.uint32)
    first, _ = km.compress_model([a2.tolist()], 1)
    both, _ = km.compress_model([b.tolist()], 1, merge_prev_words=first)
    assert np.array_equal(np.array(both, dtype=np.uint32), orc.compress(np.concatenate([a2, b]), 1))
    # a leading run that spans the whole second segment, then a third launch continues it
    z = np.zeros(992 * 2, dtype=np.uint32)
    first, _ = km.compress_model([a.tolist()], 1)
    second, _ = km.compress_model([z.tolist()], 1, merge_prev_words=first)
    third, _ = km.compress_model([b.tolist()], 1, merge_prev_words=second)
    assert np.array_equal(np.array(third, dtype=np.uint32), orc.compress(np.concatenate([a, z, b]), 1))
    # the merged run would overflow the 30-bit counter: the previous word is saturated instead
    prev = [km.fill_word(0, km.MAX_FILL - 100)]
    both, _ = km.compress_model([b.tolist()], 1, merge_prev_words=prev)
    lead = (32 * (992 * 2 - 1)) // 31
    assert both[:2] == [km.fill_word(0, km.MAX_FILL), km.fill_word(0, lead - 100)]


def _streams():
    for name, data in _cases():
        yield name + "_b", orc.compress(data, 0), data
        yield name + "_c", orc.compress(data, 1), data
    # long fills: many output tiles per compressed word
    cw = np.array([0x80000000 | 100000, 5, 0xC0000000 | 70000, 0x80000000 | 1, 7, 0xC0000000 | 8191], dtype=np.uint32)
    yield "long_fills", cw, None
    f = lambda t, n: 0x80000000 | (t << 30) | n
    for k, tail in enumerate(([f(1, 100000)], [f(1, 8192 * 3 + 5)], [7, f(1, 8192 - 1)], [f(0, 5), f(1, 8192 * 2)], [f(1, 1)],
                              [f(0, 8192), f(1, 17)], [f(1, 8191), 3, f(1, 40)])):
        yield f"ones_tail_{k}", np.array(tail, dtype=np.uint32), None
    yield "one_fill", np.array([0x80000000 | 0x3FFFFFF], dtype=np.uint32)[:1] * 0 + np.uint32(0x80000000 | 300000), None


@pytest.mark.parametrize("name,cw,data", list(_streams()), ids=[c[0] for c in _streams()])
def test_expand_model_matches_oracle(name, cw, data):
    want = orc.decompress(cw)
    got, words, groups = km.expand_model(cw.tolist())
    assert words == want.size and groups == orc.decoded_groups(cw)
    assert np.array_equal(np.array(got, dtype=np.uint32), want)
    if data is not None:
        assert np.array_equal(want[: data.size], data)
        assert not want[data.size:].any()


@pytest.mark.parametrize("name,cw,data", list(_streams()), ids=[c[0] for c in _streams()])
def test_window_path_model_matches_oracle(name, cw, data):
    """The window path's arithmetic -- the packed offset / rank scan, the flag map, the end marker of a short tile, the
    rank of the word that covers a window's first group, the 32 -> 31 repack -- for every tile of the stream
    (wah_decompress.cu PATH_WINDOW)."""
    want = orc.decompress(cw)
    (got,) = km.expand_model_window([int(x) for x in cw])
    assert np.array_equal(np.array(got, dtype=np.uint32), want)


@pytest.mark.parametrize("n_cols,wpc", [(3, 31), (5, 992), (4, 7936 + 17), (2, 3 * 7936), (7, 100)])
def test_window_path_model_on_a_batch_of_columns(n_cols, wpc):
    """One stream holding several columns: tile k of column j starts at group j * col_groups + k * 8192 and its last
    tile holds fewer than 8192 groups (wah_decompress.cu: ColumnCursor, TileRes::tg)."""
    import datagen

    rng = np.random.default_rng(n_cols * 100 + wpc)
    cols = []
    for j in range(n_cols):
        kind = int(rng.integers(0, 4))
        cols.append([np.zeros(wpc, dtype=np.uint32), np.full(wpc, 0xFFFFFFFF, dtype=np.uint32),
                     datagen.uniform(wpc, 0.3, j), datagen.clustered(wpc, 0.4, 200, j)][kind])
    cols = np.stack(cols)
    for mode in (0, 1):
        cw, _ = orc.compress_batch(cols, mode)
        got = km.expand_model_window([int(x) for x in cw], n_cols, orc.num_groups(wpc))
        for j in range(n_cols):
            assert np.array_equal(np.array(got[j][:wpc], dtype=np.uint32), cols[j]), (mode, j)
            assert not any(got[j][wpc:])


@pytest.mark.parametrize("c_words,skip,n_bad,grid", [
    (1, 0, 0, 444), (1000, 0, 2, 444), (2048, 3, 0, 444), (2049, 0, 1, 444), (200_001, 2, 5, 444),
    (6 * 8192 + 17, 0, 3, 6),        # (a grid of 6 CTAs:) tiles of two sub-tiles, the second one row long
    (6 * 8192 * 3 - 5000, 1, 0, 6),  # three sub-tiles, a ragged last tile several rows shorter than the others
    (6 * 8192 * 2, 0, 1, 6),         # exact fit: no padding anywhere
])
def test_scan_geometry_covers_the_stream_once_and_counts_padding_out(c_words, skip, n_bad, grid):
    """The decoder's scan phase: every word fetched exactly once, tile sums add up to the stream's groups, fills of
    0 groups are counted as malformed but the scan's own padding is not (wah_decompress.cu scan_body)."""
    rng = np.random.default_rng(c_words)
    cw = np.where(rng.random(c_words) < 0.3, np.uint32(0x80000000) | rng.integers(1, 50, c_words).astype(np.uint32),
                  rng.integers(1, 0x7FFFFFFF, c_words).astype(np.uint32)).astype(np.uint32)
    bad_pos = skip + rng.choice(c_words - skip, size=min(n_bad, c_words - skip), replace=False) if n_bad else []
    cw[bad_pos] = 0xC0000000
    tiles = km.scan_geometry_model(cw, skip, grid)
    counts = np.where(cw >> 31, cw & 0x3FFFFFFF, 1).astype(np.int64)
    assert sum(t[0] for t in tiles) == int(counts[skip:].sum())
    assert sum(t[1] for t in tiles) == len(bad_pos)
    assert all(t[1] >= 0 for t in tiles)


def _scan_streams():
    """(name, stream, grid, the routes of pass 2 the stream must reach)"""
    rng = np.random.default_rng(77)
    lit = lambda: int(rng.integers(1, 0x7FFFFFFF))
    fill = lambda lo, hi: km.fill_word(int(rng.integers(0, 2)), int(rng.integers(lo, hi)))
    # 16 groups per word on average, three sub-tiles per tile: every row scanned in 32-bit arithmetic
    yield "short_fills", [fill(1, 64) if rng.random() < 0.5 else lit() for _ in range(41000)], 2, {"row_32bit"}
    # literals with a fill now and then (3 groups per word): row sums first, most rows skipped
    yield "mostly_literals", [fill(2, 60) if rng.random() < 0.07 else lit() for _ in range(20000)], 1, {"row_skipped", "row_64bit"}
    # fills of up to 2^30 - 1 groups: a warp's rows beyond 2^31 groups take the wide form; long fills write sparse entries
    huge = [fill(1, 3000) if rng.random() < 0.5 else lit() for _ in range(18000)]
    for i in (700, 710, 720, 9000, 9001, 17990):   # (three in one warp's rows, two adjacent ones, one near the end)
        huge[i] = km.fill_word(1, km.MAX_FILL)
    yield "huge_fills", huge, 1, {"row_64bit", "row_32bit"}
    # a few long fills only (one sub-tile: offsets pre-scanned, every pack looked at)
    yield "long_fills_one_subtile", [fill(100000, 400000) if i % 3 else lit() for i in range(300)], 4, {"row_64bit"}
    # every word one group: boundaries by arithmetic; the last tile is ragged
    yield "incompressible", [lit() for _ in range(12288 + 77)], 3, {"unit_tile"}
    # compressed by the oracle, both modes
    yield "oracle_block1024", orc.compress(datagen.clustered(60 * 992, 0.3, 300, 1), 0).tolist(), 2, set()
    yield "oracle_canonical", orc.compress(datagen.clustered(200 * 992, 0.001, 5000, 3), 1).tolist(), 1, set()


@pytest.mark.parametrize("name,cw,grid,must_reach", list(_scan_streams()), ids=[c[0] for c in _scan_streams()])
def test_scan_pass2_model_records_every_boundary_the_expand_phase_asks_for(name, cw, grid, must_reach):
    """Pass 2 of the decoder's scan (tiles / sub-tiles / rows / packs, the 32-bit row arithmetic, the inline
    single-boundary case, sparse entries for long fills) against the definition of the table: entry k names the word
    that holds group 1024 k, and that word's group offset."""
    import bisect
    cw = [int(w) for w in cw]
    cnt = [km.word_groups(w) for w in cw]
    off = [0]
    for c_ in cnt:
        off.append(off[-1] + c_)
    for ct in ((8,) if name == "huge_fills" else (1, 8)):   # (a table entry per tile of a 2^30-group fill is a million entries)
        routes = {}
        got, _ = km.scan_pass2_model(cw, grid=grid, chunk_tiles=ct, routes=routes)
        assert must_reach <= set(routes), routes
        for k, (wi, woff) in got.items():
            i = bisect.bisect_right(off, 1024 * k) - 1          # the last word that starts at or before the boundary
            assert i < len(cw) and (wi, woff) == (i, off[i]) and off[i] <= 1024 * k < off[i + 1], (k, wi, woff, i)
        # what must be there: every boundary of a word that holds up to four; of a longer fill the nine at either end and
        # the chunk starts in between
        for i, c_ in enumerate(cnt):
            k_first, k_end = (off[i] + 1023) >> 10, (off[i] + c_ + 1023) >> 10
            if k_end - k_first <= 4:
                need = range(k_first, k_end)
            else:
                need = list(range(k_first, k_first + 9)) + list(range(k_end - 9, k_end)) + list(range((k_first + ct - 1) // ct * ct, k_end, ct))
            for k in need:
                assert k in got, (i, k, k_first, k_end)
        # ... and that is all the expand phase asks for (expand_body, "the chunk's entries must have been recorded"): per
        # chunk of ct tiles the entry of its first tile and, unless the fill word that entry names reaches behind the
        # chunk's end, the entries of its other tiles and of the tile behind it -- as far as they lie inside the stream
        G = off[-1]
        out_tiles = (G + 1023) >> 10
        for k0 in range(0, out_tiles, ct):
            nt = min(ct, out_tiles - k0)
            assert k0 in got, k0
            wi, woff = got[k0]
            if (cw[wi] & km.BIT31) and woff + cnt[wi] > ((k0 + nt) << 10):
                continue
            for k in range(k0, k0 + nt + 1):
                assert k in got or (k << 10) >= G, (k0, k, wi, woff, cnt[wi])


def test_scan_geometry_recognises_unit_tiles():
    cw = np.arange(1, 5000, dtype=np.uint32)          # literals only
    assert all(t[2] for t in km.scan_geometry_model(cw))
    cw[100] = 0x80000002
    assert not km.scan_geometry_model(cw)[0][2]
