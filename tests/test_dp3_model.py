"""CPU: the skewed-lane formulation of the MAS recurrence (tests/dp3_model.py, transcribed into
csrc/mas_dp3.cuh) against the oracle -- paths, scores and the direction words bit for bit."""
import numpy as np
import pytest

import oracle
from conftest import rect_mask
from dp3_model import maximum_path


def cases():
    rng = np.random.default_rng(11)
    out = [(1, 1), (1, 40), (5, 5), (31, 64), (32, 33), (33, 65), (64, 64), (65, 97), (70, 300), (128, 129),
           (190, 200), (3, 97), (100, 129), (256, 300)]
    for _ in range(6):
        tx = int(rng.integers(1, 140))
        out.append((tx, int(rng.integers(tx, tx + 200))))
    return out


@pytest.mark.parametrize("kind", ["ties", "normal", "logdensity"])
def test_skewed_lane_model_matches_oracle(kind):
    rng = np.random.default_rng(5)
    for k, (tx, ty) in enumerate(cases()):
        if kind == "logdensity":  # the fused kernels' values: log-densities, all negative
            value = (-rng.random((tx, ty)) * 100 - 50).astype(np.float32)
        elif kind == "ties":      # integer scores incl. exact zeros: many ties (strict '<' of core.pyx:34)
            value = rng.integers(-3, 2, (tx, ty)).astype(np.float32)
        else:
            value = (rng.standard_normal((tx, ty)) * 4).astype(np.float32)
        want, wscore = oracle.maximum_path(value[None], rect_mask([tx], [ty], tx, ty), return_scores=True)
        got, score, info = maximum_path(value, tx, ty)
        assert np.array_equal(got, want[0].astype(np.int32)), (tx, ty, kind)
        assert score == wscore[0], (tx, ty, kind)
        if info["nblk"] > 4:
            assert info["guard_blocks"] < info["nblk"]      # the middle blocks really ran unguarded
