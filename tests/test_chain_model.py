"""CPU: the cluster formulation of MAS (tests/chain_model.py: tile ranges per CTA, boundary exchange with its seed
value, windowed backtrack with hand-over) reproduces the oracle's durations and score bit for bit."""
import numpy as np
import pytest

import oracle
from chain_model import maximum_path_cluster


@pytest.mark.parametrize("cs,xs", [(4, 128), (2, 256)])
def test_cluster_formulation_matches_oracle(cs, xs):
    rng = np.random.default_rng(10 * cs)
    shapes = [(512, 700), (257, 257), (300, 340), (385, 640), (129, 500), (130, 131), (64, 200), (1, 40), (511, 600)]
    for tx, ty in shapes:
        value = (rng.standard_normal((512, ty)) * 3 - 20).astype(np.float32)
        mask = np.zeros((1, 512, ty), np.float32)
        mask[0, :tx, :ty] = 1
        want, wsc = oracle.maximum_path(value[None] * mask, mask, return_scores=True)
        dur, score = maximum_path_cluster(value, tx, ty, cs, xs)
        assert np.array_equal(dur, want[0].sum(-1).astype(np.int32)), (tx, ty)
        assert score == wsc[0], (tx, ty)
