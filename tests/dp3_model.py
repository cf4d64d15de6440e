"""Lane-by-lane numpy model of the skewed-lane MAS recurrence of csrc/mas_dp3.cuh (test infrastructure).

The CUDA kernels cannot run in the build container, so the index logic of the new formulation --
which lane touches which frame at which iteration, how the direction words are re-aligned, when the
guards are needed, what crosses a warp boundary -- is stated here once in plain Python, executed
warp-by-warp exactly as the device code does, and checked against the oracle (and through it
against the reference, core.pyx:17-35) in tests/test_dp3_model.py.  The device code is a
transcription of this file.

Formulation (one utterance, 1 <= t_x <= t_y):
  * W warps, NL = 32 W lanes; global lane G = 32 w + r owns the XPL = ceil(t_x / NL) consecutive
    tokens x = G*XPL + j.
  * warp-local iteration tau: lane r works on frame f = tau - r (one frame behind its left
    neighbour).  The only cross-lane value, V[x-1, f-1] for the lane's first token, was produced by
    lane r-1 TWO iterations earlier, so its shuffle is issued a whole iteration before it is used:
    the recurrence never waits for a shuffle.
  * direction bit of iteration tau goes to bit tau & 31 of `acc`; at the end of every 32-iteration
    block the frame-aligned word of chunk g-1 is funnelshift_r(prev, acc, r).
  * warp w+1's lane 0 reads V[last token of warp w][f-1] from a ring that warp w's lane 31 fills.
  * GUARD blocks apply "x > f -> -1e9" (head: also covers f < 0) and "f >= t_y -> keep" (tail);
    BODY blocks apply nothing.  mode "select" (drop-in, bit-exact on any input): head guard while
    any lane still has x > f.  mode "fmax" (fused: values are log-densities, all negative): head
    guard in block 0 only -- cells with x > f then drift below -1e9 and can never win a max
    against a reachable cell.
"""
from __future__ import annotations

import numpy as np

NEG = np.float32(-1e9)


def funnelshift_r(lo, hi, r):
    r &= 31
    v = (int(hi) << 32) | int(lo)
    return np.uint32((v >> r) & 0xFFFFFFFF)


def dp3_forward(value, tx, ty, W=2, mode="select", poison=True):
    """value [rows >= tx, cols >= ty] fp32.  Returns (bits [nch, NL*XPL] uint32, score, info)."""
    value = np.asarray(value, np.float32)
    NL = 32 * W
    XPL = max(1, -(-tx // NL))
    nch = -(-ty // 32)
    nrows = NL * XPL
    bits = np.zeros((nch, nrows), np.uint32)
    nblk = -(-(ty + 31) // 32)
    edge = [dict() for _ in range(W)]
    score = None
    guard_blocks = 0

    def v_of(x, f):
        if 0 <= f < ty and x < tx:
            return value[x, f]
        return np.float32(np.nan) if poison else np.float32(0)   # stale shared memory

    for w in range(W):
        V = np.full((32, XPL), NEG, np.float32)
        acc = np.zeros((32, XPL), np.uint32)
        prev = np.zeros((32, XPL), np.uint32)
        cap = np.full(32, NEG, np.float32)          # shuffle issued at the top of the previous iteration
        for g in range(nblk):
            x_last = (32 * w + 32) * XPL - 1        # largest token of lane 31
            head = (32 * g - 31 < x_last) if mode == "select" else (g == 0)
            tail = 32 * g + 31 >= ty
            guard = head or tail
            guard_blocks += guard
            for i in range(32):
                tau = 32 * g + i
                cap_new = np.empty(32, np.float32)
                cap_new[1:] = V[:-1, XPL - 1]       # __shfl_up_sync(V[XPL-1], 1): state at the end of tau-1
                cap_new[0] = NEG
                newV = V.copy()
                for r in range(32):
                    G = 32 * w + r
                    f = tau - r
                    if r > 0:
                        left = cap[r]
                    elif w == 0:
                        left = np.float32(0.0) if f == 0 else NEG
                    else:
                        left = NEG if f <= 0 else edge[w - 1].get(f - 1, np.float32(np.nan))
                    for j in range(XPL - 1, -1, -1):
                        x = G * XPL + j
                        up = left if j == 0 else V[r, j - 1]
                        cur = V[r, j]
                        with np.errstate(invalid="ignore"):
                            take = bool(up > cur)
                            if mode == "select":
                                m = up if take else cur
                            else:
                                m = np.fmax(up, cur)
                            nv = np.float32(m + v_of(x, f))
                        if take:
                            acc[r, j] |= np.uint32(1 << (tau & 31))
                        if guard:
                            if f >= ty:
                                nv = cur
                            elif x > f:
                                nv = NEG
                        newV[r, j] = nv
                    if r == 31 and w < W - 1:
                        edge[w][f] = newV[r, XPL - 1]
                V = newV
                cap = cap_new
            # ---- end of block g: frame-aligned words of chunk g-1
            c = g - 1
            for r in range(32):
                G = 32 * w + r
                for j in range(XPL):
                    x = G * XPL + j
                    word = funnelshift_r(prev[r, j], acc[r, j], r)
                    if 0 <= c < nch:
                        if (x >> 5) == c:
                            word |= np.uint32(1 << (x & 31))     # x == y always steps down (core.pyx:34)
                        if x == 0:
                            word = np.uint32(0)                    # token 0 never does
                        bits[c, x] = word
                    prev[r, j] = acc[r, j]
                    acc[r, j] = 0
        c = nblk - 1                                  # final flush (t_y = 1 mod 32)
        if c < nch:
            for r in range(32):
                G = 32 * w + r
                for j in range(XPL):
                    x = G * XPL + j
                    word = funnelshift_r(prev[r, j], 0, r)
                    if (x >> 5) == c:
                        word |= np.uint32(1 << (x & 31))
                    if x == 0:
                        word = np.uint32(0)
                    bits[c, x] = word
        ql, qj = divmod(tx - 1, XPL)
        if ql // 32 == w:
            score = V[ql % 32, qj]
    return bits, score, {"XPL": XPL, "nblk": nblk, "guard_blocks": guard_blocks}


def backtrack(bits, tx, ty):
    """core.pyx:32-35 on the direction words: idx -= d[idx, y]."""
    path = np.zeros((tx, ty), np.int32)
    idx = tx - 1
    for y in range(ty - 1, -1, -1):
        path[idx, y] = 1
        if idx != 0 and (int(bits[y >> 5, idx]) >> (y & 31)) & 1:
            idx -= 1
    return path


def maximum_path(value, tx, ty, W=2, mode="select"):
    bits, score, info = dp3_forward(value, tx, ty, W, mode)
    return backtrack(bits, tx, ty), score, info
