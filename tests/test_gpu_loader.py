"""Loader standardisation on the device against the oracle (itself pinned to scikit-learn on the CPU)."""
import numpy as np
import pytest
import torch

from oracle.loader import standardize

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


@pytest.mark.parametrize("n,c,skip_first", [(5000, 29, True), (20000, 20, False), (7, 4, False), (3000, 130, True)])
def test_standardize_matches_oracle(n, c, skip_first):
    from dgnn_b200.data import standardize_
    rng = np.random.default_rng(n + c)
    x = (rng.standard_normal((n, c)) * rng.uniform(0.01, 50, c) + rng.uniform(-20, 20, c)).astype(np.float32)
    if c > 6:
        x[:, 5] = 3.25
        x[:, 6] = 0.0
    ref = standardize(x, skip_first=skip_first)
    out = standardize_(torch.from_numpy(x).to(DEV), skip_first=skip_first).cpu().numpy()
    np.testing.assert_allclose(out, ref, rtol=2e-6, atol=2e-6)
    if skip_first:
        assert np.array_equal(out[:, 0], x[:, 0])
    # run-to-run reproducible (fixed-order reductions)
    out2 = standardize_(torch.from_numpy(x).to(DEV), skip_first=skip_first).cpu().numpy()
    assert np.array_equal(out, out2)
