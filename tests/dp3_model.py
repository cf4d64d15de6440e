"""Lane-by-lane numpy model of the skewed-lane MAS recurrence of csrc/mas_dp3.cuh (test infrastructure).

The CUDA kernels cannot run in the build container, so the index logic of the new formulation --
which lane touches which frame at which iteration, how the direction words are re-aligned, when the
guards are needed, what crosses a warp boundary -- is stated here once in plain Python, executed
warp-by-warp exactly as the device code does, and checked against the oracle (and through it
against the reference, core.pyx:17-35) in tests/test_dp3_model.py.  The device code is a
transcription of this file.

Formulation (one utterance, 1 <= t_x <= t_y, one warp): see dp3_forward below and the header of
csrc/mas_dp3.cuh.
"""
from __future__ import annotations

import numpy as np

NEG = np.float32(-1e9)


def funnelshift_r(lo, hi, r):
    r &= 31
    v = (int(hi) << 32) | int(lo)
    return np.uint32((v >> r) & 0xFFFFFFFF)


def brev(v):
    return np.uint32(int(f"{int(v):032b}"[::-1], 2))


def dp3_forward(value, tx, ty, W=1, mode="diffsign", poison=True):
    """value [rows >= tx, cols >= ty] fp32.  Returns (bits [nch, 32*XPL] uint32, score, info).

    The formulation of csrc/mas_dp3.cuh, ONE warp (W is accepted for the tests' signature only):
      cell:   d = cur - up;  acc = (acc << 1) | signbit(d);  V = fmax(up, cur) + v
              (up > cur  <=>  cur - up < 0 in IEEE arithmetic without NaN; V is never -0)
      block 0 is guarded ("x > f -> -1e9", which also covers f < 0); later blocks are not: the
      PRODUCER stores 0.0 for every cell above the diagonal (x > f), so those cells stay exactly
      -1e9 (max(-1e9, -1e9) + 0) without a select; the last blocks capture V when f == ty - 1
      (the score) instead of freezing the recurrence; direction words are accumulated MSB first
      and bit-reversed when they are re-aligned."""
    assert W == 1
    value = np.asarray(value, np.float32)
    XPL = max(1, -(-tx // 32))
    nch = -(-ty // 32)
    bits = np.zeros((nch, 32 * XPL), np.uint32)
    nblk = -(-(ty + 31) // 32)
    guard_blocks = 0

    def v_of(x, f):
        if x < tx and 0 <= f < ty:
            return value[x, f] if x <= f else np.float32(0.0)      # producer: zeros above the diagonal
        if x < tx and f < 0:
            return np.float32(np.nan) if poison else np.float32(0)  # stale tile: only block 0 reads it
        return np.float32(np.nan) if poison else np.float32(0)      # f >= ty or padding rows: garbage

    V = np.full((32, XPL), NEG, np.float32)
    acc = np.zeros((32, XPL), np.uint32)
    prev = np.zeros((32, XPL), np.uint32)       # previous block, already bit-reversed (LSB = first iteration)
    sv = np.zeros((32, XPL), np.float32)
    cap = np.full(32, NEG, np.float32)
    for g in range(nblk):
        head = g == 0
        tail = 32 * g + 31 >= ty - 1
        guard_blocks += head or tail
        for i in range(32):
            tau = 32 * g + i
            cap_new = np.empty(32, np.float32)
            cap_new[1:] = V[:-1, XPL - 1]
            cap_new[0] = NEG
            newV = V.copy()
            for r in range(32):
                f = tau - r
                left = cap[r] if r > 0 else (np.float32(0.0) if f == 0 else NEG)
                for j in range(XPL - 1, -1, -1):
                    x = r * XPL + j
                    up = left if j == 0 else V[r, j - 1]
                    cur = V[r, j]
                    with np.errstate(invalid="ignore", over="ignore"):
                        d = np.float32(cur - up)
                        nv = np.float32(np.fmax(up, cur) + v_of(x, f))
                    acc[r, j] = np.uint32(((int(acc[r, j]) << 1) & 0xFFFFFFFF) | int(np.signbit(d)))
                    if head and x > f:
                        nv = NEG
                    newV[r, j] = nv
                    if tail and f == ty - 1:
                        sv[r, j] = nv
            V = newV
            cap = cap_new
        c = g - 1
        for r in range(32):
            for j in range(XPL):
                x = r * XPL + j
                cur_n = brev(acc[r, j])
                word = funnelshift_r(prev[r, j], cur_n, r)
                if 0 <= c < nch:
                    if (x >> 5) == c:
                        word |= np.uint32(1 << (x & 31))     # x == y always steps down (core.pyx:34)
                    if x == 0:
                        word = np.uint32(0)                    # token 0 never does
                    bits[c, x] = word
                prev[r, j] = cur_n
                acc[r, j] = 0
    c = nblk - 1                                  # final flush (t_y = 1 mod 32)
    if c < nch:
        for r in range(32):
            for j in range(XPL):
                x = r * XPL + j
                word = funnelshift_r(prev[r, j], 0, r)
                if (x >> 5) == c:
                    word |= np.uint32(1 << (x & 31))
                if x == 0:
                    word = np.uint32(0)
                bits[c, x] = word
    ql, qj = divmod(tx - 1, XPL)
    return bits, sv[ql, qj], {"XPL": XPL, "nblk": nblk, "guard_blocks": guard_blocks}


def backtrack(bits, tx, ty):
    """core.pyx:32-35 on the direction words: idx -= d[idx, y]."""
    path = np.zeros((tx, ty), np.int32)
    idx = tx - 1
    for y in range(ty - 1, -1, -1):
        path[idx, y] = 1
        if idx != 0 and (int(bits[y >> 5, idx]) >> (y & 31)) & 1:
            idx -= 1
    return path


def maximum_path(value, tx, ty, W=1, mode="diffsign"):
    bits, score, info = dp3_forward(value, tx, ty, W, mode)
    return backtrack(bits, tx, ty), score, info
