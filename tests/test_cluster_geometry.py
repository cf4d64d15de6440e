"""CPU: the tile / boundary-exchange arithmetic of the cluster kernel (art_tts_b200/csrc/mas_prior_tc.cu `geometry`,
the in_last / out_first of the DP role, mas_dp.cuh dp_forward_chain), restated in Python and checked by brute force
against the band of core.pyx:18 for every CTA of a cluster."""
import numpy as np
import pytest


def cta_tiles(tx, ty, xlo, xs):
    """(t_lo, ntiles) of the CTA owning tokens [xlo, xlo + xs): what `geometry` computes."""
    nt = (ty + 31) // 32
    if tx <= xlo:
        return 0, 0
    t_lo = xlo >> 5
    t_hi = min(nt - 1, (min(tx, xlo + xs) - 1 - tx + ty) >> 5)
    return t_lo, t_hi - t_lo + 1


def in_band(x, y, tx, ty):
    return max(0, tx + y - ty) <= x <= min(tx - 1, y)


@pytest.mark.parametrize("cs,xs", [(4, 128), (2, 256)])
def test_every_band_cell_is_in_its_ctas_tiles_and_every_boundary_value_is_exchanged(cs, xs):
    rng = np.random.default_rng(cs)
    cases = [(512, 4096), (512, 512), (257, 257), (300, 301), (385, 1000), (129, 4000), (511, 600)]
    cases += [(int(a), int(a + b)) for a, b in zip(rng.integers(1, 513, 40), rng.integers(0, 1500, 40))]
    for tx, ty in cases:
        nt = (ty + 31) // 32
        for h in range(cs):
            xlo = h * xs
            t_lo, ntiles = cta_tiles(tx, ty, xlo, xs)
            xhi = min(tx, xlo + xs) - 1
            # (1) the tiles that hold a band cell of this CTA's tokens are exactly [t_lo, t_lo + ntiles)
            y = np.arange(ty)
            blo, bhi = np.maximum(0, tx + y - ty), np.minimum(tx - 1, y)
            hit = (blo <= bhi) & (blo <= xhi) & (bhi >= xlo) if xlo <= xhi else np.zeros(ty, bool)
            need = sorted(set((y[hit] >> 5).tolist()))
            assert need == list(range(t_lo, t_lo + ntiles)), (tx, ty, h, need[:3], need[-3:], t_lo, ntiles)
            # (2) the link to the right neighbour: every band cell of its first token needs V[xlo'-1, y-1] from a
            #     tile this CTA both works on and publishes (tiles out_first .. t_hi; out_first = t_lo' - 1)
            xr = xlo + xs
            if h < cs - 1 and tx > xr:
                out_first = (xr >> 5) - 1
                t_hi = t_lo + ntiles - 1
                r_lo, r_n = cta_tiles(tx, ty, xr, xs)
                in_last = min(nt - 1, (xr - 1 - tx + ty) >> 5)          # what the right CTA's warp 0 computes
                assert in_last == t_hi and out_first == r_lo - 1 and out_first >= t_lo
                for y in range(1, ty):
                    if in_band(xr, y, tx, ty):                           # the diagonal predecessor is in the band too
                        assert in_band(xr - 1, y - 1, tx, ty)
                        tprev = (y - 1) >> 5
                        assert out_first <= tprev <= t_hi, (tx, ty, h, y)
                        # and the right CTA is working on the tile of frame y, no earlier than one tile after the seed
                        assert r_lo <= (y >> 5) <= r_lo + r_n - 1
