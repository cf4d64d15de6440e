"""numpy model of the CLUSTER formulation of MAS (art_tts_b200/csrc/mas_dp.cuh dp_forward_chain +
backtrack_bits_window, mas_prior_tc.cu `geometry`): the token axis of one utterance split over `cs` CTAs of `xs`
tokens, every CTA working only on the 32-frame tiles that hold band cells of its tokens, the recurrence crossing
the CTA boundary through a ring of boundary values (one per frame), the backtrack handed from CTA to CTA.
Test infrastructure (tests/test_chain_model.py compares it with the oracle); not a product path."""
import numpy as np

NEG = np.float32(-1e9)


def cta_tiles(tx, ty, xlo, xs):
    nt = (ty + 31) // 32
    if tx <= xlo:
        return 0, 0
    t_lo = xlo >> 5
    return t_lo, min(nt - 1, (min(tx, xlo + xs) - 1 - tx + ty) >> 5) - t_lo + 1


def forward_cta(value, tx, ty, xlo, xs, edge_in):
    """Tokens [xlo, min(tx, xlo+xs)) over the CTA's tiles.  edge_in[y] = V[xlo-1, y] for the frames the left CTA
    published (NaN elsewhere).  Returns (bits [nt, xs] uint32, edge_out [ty] fp32 with NaN where nothing was
    published, score of token tx-1 or None, first tile, tiles)."""
    nt = (ty + 31) // 32
    t_lo, ntiles = cta_tiles(tx, ty, xlo, xs)
    n = min(tx, xlo + xs) - xlo
    bits = np.zeros((nt, xs), np.uint32)
    edge_out = np.full(ty, np.nan, np.float32)
    if ntiles == 0:
        return bits, edge_out, None, t_lo, ntiles
    xg = xlo + np.arange(n)                       # global token indices
    V = np.full(n, NEG, np.float32)
    # `left` of the first local token before the first frame of the first tile
    if xlo == 0:
        left = np.float32(0.0)
    else:
        y_seed = 32 * t_lo - 1                    # last frame of the left CTA's tile t_lo - 1
        assert not np.isnan(edge_in[y_seed]), "the seed value was not published"
        left = edge_in[y_seed]
    for t in range(t_lo, t_lo + ntiles):
        for y in range(32 * t, min(ty, 32 * t + 32)):
            up = np.empty(n, np.float32)
            up[1:] = V[:-1]
            up[0] = left
            take = up > V
            d = (xg != 0) & ((xg == y) | take)
            new = (np.where(take, up, V) + value[xg, y]).astype(np.float32)
            V = np.where(xg <= y, new, NEG).astype(np.float32)
            bits[t, :n] |= d.astype(np.uint32) << np.uint32(y & 31)
            edge_out[y] = V[n - 1]                # what warp 1's last lane publishes
            # value of token xlo-1 at THIS frame is the `left` of the next frame
            if xlo == 0:
                left = NEG
            else:
                left = edge_in[y] if not np.isnan(edge_in[y]) else NEG     # outside the exchange range: don't care
    score = V[tx - 1 - xlo] if xlo <= tx - 1 < xlo + xs else None
    return bits, edge_out, score, t_lo, ntiles


def backtrack_cta(bits, n, xlo, idx, y, first, dur, win_c=8):
    """The windowed walk of backtrack_bits_window over the CTA's local tokens (window = 32 tokens x win_c chunks,
    reloaded when the walk leaves it).  Returns the frame handed to the left CTA, or -1."""
    top, c = y, y >> 5
    wx = wc = -1
    reloads = 0
    while y >= 0 and idx != 0:
        xl = idx - xlo
        if wx < 0 or xl < wx - 31 or c < wc - (win_c - 1):
            wx, wc = xl, c
            reloads += 1
        assert wx - 31 <= xl <= wx and wc - (win_c - 1) <= c <= wc
        w = int(bits[c, xl])
        m = w & (0xffffffff >> (31 - (y & 31)))
        if m == 0:
            y = (c << 5) - 1
            c -= 1
            continue
        p = m.bit_length() - 1
        ys = (c << 5) + p
        first[idx], dur[idx] = ys, top - ys + 1
        idx -= 1
        y = ys - 1
        top = y
        if idx < xlo:
            return y
        if p == 0:
            c -= 1
    if top >= 0:
        first[idx], dur[idx] = 0, top + 1
    return -1


def maximum_path_cluster(value, tx, ty, cs, xs):
    """Durations [T_x] of one utterance (1 <= tx <= ty) by the cluster formulation."""
    T_x = value.shape[0]
    edge = np.full(ty, np.nan, np.float32)
    per_cta = []
    score = None
    for h in range(cs):
        xlo = h * xs
        # the left CTA publishes only the tiles from (xlo >> 5) - 1 on (out_first); mask the rest to prove that
        if h > 0:
            e = edge.copy()
            e[: max(0, 32 * ((xlo >> 5) - 1))] = np.nan
        else:
            e = edge
        bits, edge, sc, t_lo, ntiles = forward_cta(value, tx, ty, xlo, xs, e)
        per_cta.append((bits, xlo, ntiles))
        if sc is not None:
            score = sc
    first = np.zeros(T_x, np.int64)
    dur = np.zeros(T_x, np.int64)
    y = ty - 1
    for h in range(cs - 1, -1, -1):
        bits, xlo, ntiles = per_cta[h]
        if ntiles == 0:
            continue
        n = min(tx, xlo + xs) - xlo
        y = backtrack_cta(bits, n, xlo, xlo + n - 1, y, first, dur)
        if xlo > 0:
            assert y >= xlo - 1, "a token never ends before its own index"
    return dur.astype(np.int32), score
