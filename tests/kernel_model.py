"""Executable model of the arithmetic inside the CUDA kernels (pure Python, tiny inputs only).

There is no GPU on the development box, so the index arithmetic of ``gpu-wah_b200/csrc/wah_compress.cu`` and
``wah_decompress.cu`` -- tail masks, the open-run carry through thread / warp / tile aggregates, the aggregate
monoid over packed tile descriptors (the kernel now sums the descriptors between a CTA's consecutive tiles; the
monoid and its result are the same as for the look-back walk modelled here), the seam between chained launches,
output-tile boundary bookkeeping, the word-centric expansion -- is restated here and checked against the oracle in
``test_kernel_model.py``.  This is TEST INFRASTRUCTURE: it is not a fallback and nothing in the product imports it.
"""
from __future__ import annotations


ONES31 = 0x7FFFFFFF
BIT31 = 0x80000000
BIT30 = 0x40000000
MAX_FILL = 0x3FFFFFFF
M32 = 0xFFFFFFFF


def popc(x):
    return bin(x & M32).count("1")


def clz(x):
    x &= M32
    return 32 - x.bit_length()


def ffs(x):
    x &= M32
    return (x & -x).bit_length()  # 1-based, 0 if none


def funnelshift_r(lo, hi, s):
    return (((hi << 32) | lo) >> (s & 31)) & M32


def extract_group(row, g):
    bit = 31 * g
    wi, s = bit >> 5, bit & 31
    return funnelshift_r(row[wi], row[wi + 1], s) & ONES31


def fill_word(t, n):
    return BIT31 | (t << 30) | n


def word_groups(w):
    return (w & MAX_FILL) if (w & BIT31) else 1


# ------------------------------------------------------------------ compress model

LBN = 10   # descriptors per lane and round of the control warps' window (wah_compress.cu)


def chained_offsets(aggs, grid, tiles_per_col, block_mode):
    """Mirror of the control warps of wah_compress_kernel: the chained sum that gives every tile the number of words
    the launch emits before it (excl) and the length of the run still open where it starts (carry).

    aggs[k] = (tile_cnt, tile_open, tile_has) as the workers publish them: words the tile emits, groups behind its
    last run end (the whole tile if it has none), whether it holds a run end.  CTA b of `grid` owns tiles b, b + grid,
    ...; its two control warps take them in turn, the warp of tile i handing (cnt_end, open_end) to the warp of tile
    i + 1 (sm.chain).  A warp sums the descriptors of the tiles between the CTA's previous tile and this one, nearest
    first, 32 lanes x LBN descriptors per round; towards older tiles the open groups add up until a tile with a run
    end is met."""
    n_tiles = len(aggs)
    out = [None] * n_tiles
    for b in range(min(grid, n_tiles)):
        chain = None   # (cnt_end, open_end) of the CTA's previous tile, written by the OTHER control warp
        for i, tile in enumerate(range(b, n_tiles, grid)):
            t = tile % tiles_per_col
            lo = 0 if i == 0 else tile - grid + 1
            hi = tile - 1
            open_done = block_mode or t == 0
            lane_csum, lane_osum = [0] * 32, [0] * 32
            while hi >= lo:
                for r in range(LBN):
                    lk = [hi - lane - 32 * r for lane in range(32)]
                    inside = [k >= lo for k in lk]
                    for lane in range(32):
                        if inside[lane]:
                            lane_csum[lane] += aggs[lk[lane]][0]
                    if not open_done:
                        term = [inside[lane] and aggs[lk[lane]][2] for lane in range(32)]
                        first_term = next((lane for lane in range(32) if term[lane]), 32)
                        for lane in range(32):
                            if inside[lane] and lane <= first_term:
                                lane_osum[lane] += aggs[lk[lane]][1]
                        open_done = first_term < 32
                hi -= 32 * LBN
            between, open_between = sum(lane_csum), sum(lane_osum)
            own_cnt, own_open = chain if chain is not None else (0, 0)
            excl = own_cnt + between
            carry = open_between + (0 if open_done else own_open)
            if block_mode or t == 0:
                carry = 0
            tile_cnt, tile_open, tile_has = aggs[tile]
            chain = (excl + tile_cnt, tile_open if tile_has else carry + tile_open)
            out[tile] = (excl, carry)
    return out


def seam_model(seg, groups, out):
    """wah_lead_probe_kernel + wah_seam_kernel (wah_compress.cu): how the leading run of the next launch's
    segment joins the last word written so far.  May take that word back (out.pop()) or saturate it; returns
    the amount the next launch adds to its first word."""
    base, adj = len(out), 0
    if base > 0 and seg:
        typ = seg[0] & 1
        pattern = M32 if typ else 0
        bits = 32 * len(seg)
        for i, w in enumerate(seg):
            x = (w ^ pattern) & M32
            if x:
                bits = 32 * i + ffs(x) - 1
                break
        lead = groups if (bits == 32 * len(seg) and typ == 0) else bits // 31
        pw = out[base - 1]
        if lead > 0 and (pw & BIT31) and ((pw >> 30) & 1) == typ:
            c1 = pw & MAX_FILL
            if c1 + lead <= MAX_FILL:
                out.pop()
                adj = c1
            else:
                out[base - 1] = fill_word(typ, MAX_FILL)
                adj = c1 - MAX_FILL
    return adj


def compress_model(cols, mode, threads=256, merge_prev_words=None, seed=0, grid=3):
    """cols: list of equally long word lists (one launch).  mode 0 = BLOCK1024, 1 = CANONICAL.
    merge_prev_words: output of the earlier launches of the same stream (chained launches).
    Returns (out, col_offsets)."""
    block_mode = mode == 0
    NW = threads // 32
    TW, TG = threads * 31, threads * 32
    n_words = len(cols[0])
    groups = (32 * n_words + 30) // 31
    tiles_per_col = (n_words + TW - 1) // TW
    n_tiles = tiles_per_col * len(cols)
    out = list(merge_prev_words) if merge_prev_words else []
    lead_adjust = seam_model(cols[0], groups, out) if (merge_prev_words is not None and mode == 1) else 0
    base = len(out)
    aggs, classified = [], []
    col_offsets = [0] * (len(cols) + 1)

    for tile in range(n_tiles):
        col, t = divmod(tile, tiles_per_col)
        w0 = t * TW
        src = cols[col]
        left = n_words - w0
        nload = TW + 1 if left > TW else left
        s_in = [src[w0 + i] if i < nload else 0 for i in range(TW + 4)]
        th = []
        for tid in range(threads):
            lane = tid & 31
            row = s_in[31 * tid: 31 * tid + 33]
            g_thread = t * TG + 32 * tid
            nvalid = 0
            if g_thread < groups:
                nvalid = min(32, groups - g_thread)
            vmask = M32 if nvalid == 32 else ((1 << nvalid) - 1)
            Z = O = 0
            prev = row[0]
            u = (prev << 1) & M32          # u = group << 1 | one junk bit
            Z |= 1 if (u & 0xFFFFFFFE) == 0 else 0
            O |= 1 if (~u & 0xFFFFFFFE) == 0 else 0
            for j in range(1, 31):
                cur = row[j]
                u = funnelshift_r(prev, cur, 31 - j)
                Z |= (1 << j) if (u & 0xFFFFFFFE) == 0 else 0
                O |= (1 << j) if (~u & 0xFFFFFFFE) == 0 else 0
                prev = cur
            u = prev
            Z |= BIT31 if (u & 0xFFFFFFFE) == 0 else 0
            O |= BIT31 if (~u & 0xFFFFFFFE) == 0 else 0
            Z &= vmask
            O &= vmask
            F = Z | O
            nz = no = 0
            if not (block_mode and lane == 31) and g_thread + 32 < groups:
                nx = row[31] & ONES31
                nz = BIT31 if nx == 0 else 0
                no = BIT31 if nx == ONES31 else 0
            T = ((~F) & vmask) | (Z & ~((Z >> 1) | nz)) | (O & ~((O >> 1) | no))
            T &= M32
            th.append(dict(Z=Z, O=O, F=F, T=T, cnt=popc(T), my_open=clz(T) if T else 32, row=row))
        s_wcnt, s_wopen, s_whas = [0] * NW, [0] * NW, [0] * NW
        for warp in range(NW):
            lanes = th[32 * warp: 32 * warp + 32]
            incl = 0
            tb = 0
            for lane, d in enumerate(lanes):
                incl += d["cnt"]
                d["incl"] = incl
                if d["T"]:
                    tb |= 1 << lane
            for lane, d in enumerate(lanes):
                below = tb & ((1 << lane) - 1)
                q = 31 - clz(below) if below else 0
                open_q = lanes[q]["my_open"]
                d["below"] = below
                d["prev_open"] = open_q + 32 * (lane - q - 1) if below else 32 * lane
            qlast = 31 - clz(tb) if tb else 0
            s_wcnt[warp] = incl
            s_whas[warp] = 1 if tb else 0
            s_wopen[warp] = lanes[qlast]["my_open"] + 32 * (31 - qlast) if tb else 1024
        # warp 0: the tile's aggregate, published to the other CTAs as soon as the tile is classified
        tile_cnt = tile_open = tile_has = 0
        for w in range(NW):
            tile_cnt += s_wcnt[w]
            if s_whas[w]:
                tile_has, tile_open = 1, s_wopen[w]
            else:
                tile_open += s_wopen[w]
        if block_mode or t == tiles_per_col - 1:   # the end of a column (and of every block) is always a run end
            tile_open, tile_has = 0, 1
        aggs.append((tile_cnt, tile_open, tile_has))
        classified.append((s_in, th, s_wcnt, s_wopen, s_whas))

    # control warps: every tile's offset and open run from the chained sum over the published aggregates
    offsets = chained_offsets(aggs, grid, tiles_per_col, block_mode)

    for tile in range(n_tiles):
        col, t = divmod(tile, tiles_per_col)
        s_in, th, s_wcnt, s_wopen, s_whas = classified[tile]
        tile_cnt = aggs[tile][0]
        s_out = [None] * TG
        s_excl, carry = offsets[tile]
        s_carry = 0 if (block_mode or t == 0) else carry
        # emission
        for warp in range(NW):
            wprefix = wcarry = 0
            found = False
            for w in range(NW):
                if w < warp:
                    wprefix += s_wcnt[w]
                    if s_whas[w]:
                        wcarry, found = s_wopen[w], True
                    else:
                        wcarry += s_wopen[w]
            if not found:
                wcarry += s_carry
            if block_mode:
                wcarry = 0
            lanes = th[32 * warp: 32 * warp + 32]
            for d in lanes:
                if not d["below"]:
                    d["prev_open"] += wcarry
                d["my_off"] = wprefix + d["incl"] - d["cnt"]
            wcnt = s_wcnt[warp]
            wrow = s_in[992 * warp: 992 * warp + 994]
            all_literal = all(d["T"] == M32 and d["F"] == 0 for d in lanes)
            if all_literal:
                for k in range(32):
                    for lane in range(32):
                        g = 32 * k + lane
                        s_out[wprefix + g] = extract_group(wrow, g)
            elif wcnt > 192:
                for k in range(32):
                    Tk = lanes[k]["T"]
                    if Tk == 0:
                        continue
                    Fk, Ok, offk, pok = lanes[k]["F"], lanes[k]["O"], lanes[k]["my_off"], lanes[k]["prev_open"]
                    for lane in range(32):
                        bit = 1 << lane
                        if Tk & bit:
                            lower = Tk & (bit - 1)
                            if Fk & bit:
                                ln = lane - (31 - clz(lower)) if lower else lane + 1 + pok
                                w = fill_word((Ok >> lane) & 1, ln)
                            else:
                                w = extract_group(wrow, 32 * k + lane)
                            s_out[offk + popc(lower)] = w
            else:
                for d in lanes:
                    m, off, prev, extra = d["T"], d["my_off"], -1, d["prev_open"]
                    while m:
                        j = ffs(m) - 1
                        m &= m - 1
                        if (d["F"] >> j) & 1:
                            w = fill_word((d["O"] >> j) & 1, (j - prev) + extra)
                        else:
                            w = extract_group(d["row"], j)
                        s_out[off] = w
                        off += 1
                        prev, extra = j, 0
        drop = 0
        dst0 = base + s_excl
        if t == 0:
            col_offsets[col] = dst0
        if tile == n_tiles - 1:
            col_offsets[len(cols)] = dst0 + tile_cnt - drop
        need = dst0 + tile_cnt - drop
        if len(out) < need:
            out.extend([None] * (need - len(out)))
        for i in range(drop, tile_cnt):
            assert s_out[i] is not None, (tile, i)
            out[dst0 + i - drop] = s_out[i] + (lead_adjust if s_excl + i == 0 else 0)   # the launch's first word
    assert all(w is not None for w in out)
    return out, col_offsets


# ---------------------------------------------------------------- decompress model

SCAN_THREADS, SCAN_ITEMS = 256, 8
SCAN_TILE = SCAN_THREADS * SCAN_ITEMS
TG, TWO = 1024, 992


def scan_model(cw, max_out_tiles):
    """Mirror of wah_scan_kernel: returns (G, words, out_tiles, starts)."""
    c = len(cw)
    n_tiles = (c + SCAN_TILE - 1) // SCAN_TILE
    starts = {}
    base = 0
    k_limit = max_out_tiles + 1
    for tile in range(n_tiles):
        off = base
        for tid in range(SCAN_THREADS):
            w_begin = tile * SCAN_TILE + tid * SCAN_ITEMS
            for i in range(SCAN_ITEMS):
                w = cw[w_begin + i] if w_begin + i < c else BIT31
                cnt = word_groups(w)
                k_first = (off + TG - 1) // TG
                k_end = (off + cnt + TG - 1) // TG
                k_end = min(k_end, k_limit)
                k_first = min(k_first, k_end)
                for k in range(k_first, k_end):
                    assert k not in starts
                    starts[k] = (w_begin + i, off)
                off += cnt
        base = off
    G = base
    words = (G >> 5) * 31 + (((G & 31) * 31 + 31) >> 5)
    return G, words, (G + TG - 1) // TG, starts


def expand_model(cw, out_cap=None):
    """Mirror of wah_decode_kernel's expand phase for a single stream: per output tile (1024 groups = 992 words, one warp)
    the table entries of the tile and of the next one give the tile's word range; a tile inside one fill word is written
    as a constant, every other tile goes through the window path (window_tile_model).  `out_cap` = capacity of the output
    in words (the table then has room for ceil(out_cap / 992) tiles and the last tile may be cut short)."""
    c = len(cw)
    G, words, real_tiles, starts = scan_model(cw, 1 << 40 if out_cap is None else (out_cap + TWO - 1) // TWO)
    total_words = words if out_cap is None else min(words, out_cap)
    out = [None] * total_words
    n_tiles = real_tiles if out_cap is None else min(real_tiles, (out_cap + TWO - 1) // TWO)
    for ot in range(n_tiles):
        ws, g0 = starts[ot]
        last = ot + 1 >= real_tiles            # the entry of the next tile is never recorded: it lies behind the stream
        we = c - 1 if last else starts[ot + 1][0]
        g_lo = ot * TG
        skip = g_lo - g0
        tg = min(TG, G - g_lo) if last else TG
        w_lo = ot * TWO
        if w_lo >= total_words:
            continue
        nout = min(TWO, total_words - w_lo)
        first = cw[ws]
        if ws == we and (first & BIT31) and (not last or not (first & BIT30)):
            f = M32 if (first & BIT30) else 0
            for i in range(nout):
                out[w_lo + i] = f
            continue
        img = window_tile_model([int(x) for x in cw[ws:we + 1]], skip, tg)
        out[w_lo:w_lo + nout] = img[:nout]
    assert all(w is not None for w in out)
    return out, words, G


# ---------------------------------------------------------------- scan phase geometry (wah_decompress.cu scan_body)
import numpy as np  # noqa: E402

SCAN_MAXV = 8                      # 16-byte packs per lane and sub-tile
SCAN_SUB_WORDS = SCAN_MAXV * 4 * SCAN_THREADS


def scan_tile_words_model(c_words, grid=444):
    """scan_tile_words(): one tile per CTA of the decode grid, a multiple of 1024 words, at least 2048"""
    unit = 4 * SCAN_THREADS
    tw = ((c_words + grid - 1) // grid + unit - 1) // unit * unit
    return max(tw, SCAN_TILE)


def scan_geometry_model(cw, skip_words=0, grid=444):
    """Walks the stream the way scan_body does -- tiles, sub-tiles of up to 8 rows per warp, 16-byte packs fetched
    with a byte count (cp.async zero-fills the rest), padding patched to fills of 0 groups -- and returns what pass 1
    publishes per tile: (groups, malformed words, is unit tile).  Asserts that every word of the stream is fetched
    exactly once and nothing behind it is read."""
    c = len(cw)
    tw = scan_tile_words_model(c, grid)
    n_tiles = (c + tw - 1) // tw
    assert n_tiles <= grid
    seen = np.zeros(c, dtype=np.int32)
    out = []
    for tile in range(n_tiles):
        tile_begin = tile * tw
        w_first, w_last = max(tile_begin, skip_words), min(tile_begin + tw, c)
        rows = (w_last - tile_begin + 4 * SCAN_THREADS - 1) // (4 * SCAN_THREADS) if w_last - tile_begin < tw else tw // (4 * SCAN_THREADS)
        nsub = (rows + SCAN_MAXV - 1) // SCAN_MAXV
        padding = rows * 4 * SCAN_THREADS - (w_last - w_first)
        groups = zero_fills = 0
        for sub in range(nsub):
            nv = min(SCAN_MAXV, rows - sub * SCAN_MAXV)
            for warp in range(SCAN_THREADS // 32):
                seg_begin = tile_begin + sub * SCAN_SUB_WORDS + warp * nv * 128
                for v in range(nv):
                    for lane in range(32):
                        i0 = seg_begin + (v * 32 + lane) * 4
                        nbytes = 16 if i0 + 4 <= c else (max(c - i0, 0) * 4)
                        pack = [int(cw[i0 + j]) if j * 4 < nbytes else 0 for j in range(4)]     # zero filled
                        seen[i0:i0 + nbytes // 4] += 1
                        for j in range(4):                                                        # patch_sub
                            if i0 + j >= c or i0 + j < skip_words:
                                pack[j] = BIT31
                        for w in pack:
                            groups += word_groups(w)
                            zero_fills += (w & ~BIT30 & 0xFFFFFFFF) == BIT31
        out.append((groups, zero_fills - padding, w_last > w_first and groups == w_last - w_first))
    assert (seen == 1).all()
    return out


TGM = TG - 1
SCAN_SUBSUMS = 32


def _note_boundaries(starts, required, k_lim, ct, kf, ke, wi, off, pack):
    """note_boundaries / record_boundaries / write_fill_entries for a single stream: the pack's four words, the first at
    word index wi and group offset off, record the output-tile boundaries kf .. ke - 1 that fall into them.  `required`
    collects the table indices the expand phase will ask for (a long fill writes only its first and last nine entries and
    the chunk starts in between)."""
    c = [word_groups(w) for w in pack]
    if ke - kf == 1:                       # the inline case: one boundary in the pack
        if kf < k_lim:
            d = ((kf << 10) - off) & M32
            s1, s2 = (c[0] + c[1]) & M32, (c[0] + c[1] + c[2]) & M32
            jj = (c[0] <= d) + (s1 <= d) + (s2 <= d)
            before = (0, c[0], s1, s2)[jj]
            assert kf not in starts
            starts[kf] = (wi + jj, off + before)
            required.add(kf)
        return
    for j in range(4):
        if c[j]:
            k_first, k_end = (off + TGM) >> 10, min((off + c[j] + TGM) >> 10, k_lim)
            if k_first < k_end:
                ks = range(k_first, k_end)
                if k_end - k_first > 4:    # a long fill: sparse entries
                    head_end = min(k_first + 9, k_end)
                    tail_begin = k_end - 9 if k_end > head_end + 9 else head_end
                    k_mid = (head_end + ct - 1) // ct * ct
                    ks = list(range(k_first, head_end)) + list(range(tail_begin, k_end)) + list(range(k_mid, tail_begin, ct))
                for k in ks:
                    assert k not in starts or starts[k] == (wi + j, off)
                    starts[k] = (wi + j, off)
                    required.add(k)
        off += c[j]


def scan_pass2_model(cw, grid=444, max_out_tiles=1 << 40, chunk_tiles=8, routes=None):
    """Mirror of scan_body's pass 2 for a single stream (after pass 1 and the offset exchange): which compressed word
    covers every output-tile boundary k * 1024.  Follows the kernel's three methods -- boundaries by arithmetic for a
    literal-dense tile; every row scanned, in 32-bit arithmetic relative to the output-tile boundary below the warp's
    first row (the per-warp sub-tile sums kept from pass 1); row sums first and boundary-free rows skipped -- and its
    geometry (tiles, sub-tiles of 8 rows per warp, lanes of four words).  Returns (starts, required); `routes` (a dict)
    counts the rows that went each way."""
    c = len(cw)
    NW = SCAN_THREADS // 32
    tw = scan_tile_words_model(c, grid)
    n_tiles = (c + tw - 1) // tw
    k_lim = max_out_tiles + 1
    word = lambda i: int(cw[i]) if i < c else BIT31   # (behind the stream: fills of 0 groups)
    starts, required = {}, set()
    routes = {} if routes is None else routes
    count = lambda key: routes.__setitem__(key, routes.get(key, 0) + 1)
    excl = 0
    for tile in range(n_tiles):
        tile_begin, w_last = tile * tw, min(tile * tw + tw, c)
        rows = (w_last - tile_begin + 4 * SCAN_THREADS - 1) // (4 * SCAN_THREADS) if w_last - tile_begin < tw else tw // (4 * SCAN_THREADS)
        nsub = (rows + SCAN_MAXV - 1) // SCAN_MAXV
        tile_sum = sum(word_groups(word(i)) for i in range(tile_begin, tile_begin + rows * 4 * SCAN_THREADS))
        n_words_tile = w_last - tile_begin
        if tile_sum == n_words_tile and n_words_tile > 0:          # unit tile: no second look at the words
            for k in range((excl + TGM) >> 10, min((excl + tile_sum + TGM) >> 10, k_lim)):
                starts[k] = (tile_begin + (k << 10) - excl, k << 10)
                required.add(k)
            excl += tile_sum
            count("unit_tile")
            continue
        skip_rows = tile_sum < 8 * n_words_tile
        sub_base = excl
        for sub in range(nsub):
            nv = min(SCAN_MAXV, rows - sub * SCAN_MAXV)
            seg = [tile_begin + sub * SCAN_SUB_WORDS + warp * nv * 128 for warp in range(NW)]
            wsub = [sum(word_groups(word(i)) for i in range(seg[w], seg[w] + nv * 128)) for w in range(NW)]
            for warp in range(NW):
                row_base = sub_base + sum(wsub[:warp])
                narrow = wsub[warp] < (1 << 31)
                q0 = row_base & ~TGM
                rb = row_base - q0
                for v in range(nv):
                    packs = [[word(seg[warp] + (v * 32 + lane) * 4 + j) for j in range(4)] for lane in range(32)]
                    s = [sum(word_groups(w) for w in pk) for pk in packs]
                    rsum = sum(s)
                    if skip_rows and nsub > 1:
                        # (a tile of one sub-tile takes the pre-scanned route, which looks at every pack like the wide form below)
                        if ((row_base + TGM) >> 10) == ((row_base + rsum + TGM) >> 10):
                            row_base += rsum
                            count("row_skipped")
                            continue
                    e64 = row_base
                    count("row_32bit" if (narrow and not skip_rows and nsub > 1) else "row_64bit")
                    for lane in range(32):
                        wi = seg[warp] + (v * 32 + lane) * 4
                        if narrow and not skip_rows and nsub > 1:
                            e = (rb + (e64 - row_base)) & M32
                            assert e == rb + (e64 - row_base), "32-bit offset wrapped"
                            kf, ke = ((e + TGM) & M32) >> 10, ((e + s[lane] + TGM) & M32) >> 10
                            assert e + s[lane] + TGM <= M32, "32-bit boundary arithmetic wrapped"
                            if kf != ke:
                                _note_boundaries(starts, required, k_lim, chunk_tiles, (q0 >> 10) + kf, (q0 >> 10) + ke, wi, q0 + e, packs[lane])
                        else:
                            kf, ke = (e64 + TGM) >> 10, (e64 + s[lane] + TGM) >> 10
                            if kf != ke:
                                _note_boundaries(starts, required, k_lim, chunk_tiles, kf, ke, wi, e64, packs[lane])
                        e64 += s[lane]
                    row_base += rsum
                    rb += rsum
            sub_base += sum(wsub)
        excl += tile_sum
    return starts, required


# ---------------------------------------------------------------- window path of the expand phase (wah_decompress.cu)

RANK_SHIFT = 20
GROUP_MASK = (1 << RANK_SHIFT) - 1


def group_bits(x):
    if x & BIT31:
        return ONES31 if x & BIT30 else 0
    return x


def cw_pos(r):
    return r + (r >> 5)


def window_tile_model(words, skip, tg):
    """Mirror of PATH_WINDOW for ONE output tile: `words` = the compressed words ws .. we of the tile, `skip` = groups of
    the first word that belong to earlier tiles, `tg` = groups in the tile.  Step 1 (word centric): a packed scan gives
    every word its group offset (low 22 bits) and its rank among the words that hold a group (bits above); the word's
    group bits are parked by rank and a flag is set where it starts.  Step 2 (output centric): window t = groups
    32 t .. 32 t + 31 finds the rank of the word covering its first group from the number of flags below it, walks its
    32 flag bits and emits output words 31 t .. 31 t + 30.  Returns the 992-word image."""
    s_cw = [None] * (TG + TG // 32 + 8)
    s_flag = [0] * (TG // 32 + 8)
    packed = 0
    for i, wv in enumerate(words):
        c = word_groups(wv)
        if i == 0:
            c -= skip
        c = min(c, 2 * TG)
        off, rk = packed & GROUP_MASK, packed >> RANK_SHIFT
        if c != 0 and off <= tg:   # (the last word may start exactly where the tile ends: parked, but reads as zeros)
            s_cw[cw_pos(rk)] = group_bits(wv) if off < tg else 0
            s_flag[off >> 5] |= 1 << (off & 31)
        packed += c + ((1 << RANK_SHIFT) if c else 0)
        if (packed & GROUP_MASK) >= tg:   # (the kernel leaves at the end of a round of 1024 words; later words hold nothing)
            packed = (packed >> RANK_SHIFT << RANK_SHIFT) | min(packed & GROUP_MASK, GROUP_MASK)
    if tg < TG:
        s_cw[cw_pos(min(packed >> RANK_SHIFT, TG + 1))] = 0
        s_flag[tg >> 5] |= 1 << (tg & 31)
    img = [0] * TWO
    below = 0
    for t in range(TG // 32):
        F = s_flag[t]
        r = below + (F & 1) - 1
        assert r >= 0
        v = s_cw[cw_pos(r)]
        for j in range(1, 32):
            r += (F >> j) & 1
            nv = s_cw[cw_pos(r)]
            img[31 * t + j - 1] = funnelshift_r((v << 1) & M32, nv, j)
            v = nv
        below += popc(F)
    return img


def expand_model_window(cw, n_cols=1, col_groups=None):
    """The expand phase with every tile on the window path; with n_cols > 1 the stream is a batch of columns of
    col_groups groups each (output tile k of column j starts at group j * col_groups + k * 1024).  Returns the list of
    decoded columns (each a list of ceil(31 * groups / 32) words)."""
    c = len(cw)
    offs = [0]
    for w in cw:
        offs.append(offs[-1] + word_groups(w))
    G = offs[-1]
    if col_groups is None:
        col_groups = G
    assert G == n_cols * col_groups
    import bisect

    out = []
    for j in range(n_cols):
        words_out = (col_groups * 31 + 31) // 32
        col = [0] * words_out
        for k in range((col_groups + TG - 1) // TG):
            g_start = j * col_groups + k * TG
            tg = min(TG, col_groups - k * TG)
            # the word that covers group g_start: the last one that starts at or before it and holds at least a group
            ws = bisect.bisect_right(offs, g_start) - 1
            g_end = g_start + tg
            we = bisect.bisect_right(offs, g_end) - 1 if g_end < G else c - 1
            img = window_tile_model([int(x) for x in cw[ws:we + 1]], g_start - offs[ws], tg)
            nout = min(TWO, words_out - k * TWO)
            col[k * TWO:k * TWO + nout] = img[:nout]
        out.append(col)
    return out
