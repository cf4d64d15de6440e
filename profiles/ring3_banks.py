"""Brute-force check of the bank mapping of the mas_dp3.cuh ring: consumer LDS.32 of the 32 skewed
lanes and producer STS.128 (thread per row) are conflict-free for XPL = 1..8, ring of 96 or 128 frames."""
for RC in (96, 128):
    P = RC + 4
    for XPL in range(1, 9):
        for tau in range(2 * RC):
            for j in range(XPL):
                banks = {(P * (r * XPL + j) + ((tau - r) % RC)) % 32 for r in range(32)}
                assert len(banks) == 32, (RC, XPL, tau, j)
        for x0 in range(0, 32 * XPL, 8):       # quarter-warp of 8 consecutive rows, same 16-byte chunk
            for c in range(RC // 4):
                groups = {((P * x) // 4 + c) % 8 for x in range(x0, x0 + 8)}
                assert len(groups) == 8, (RC, XPL, x0, c)
print("ring3 bank mapping: conflict-free")
