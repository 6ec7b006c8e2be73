"""Randomised cross-checks of the oracle primitives whose reference implementations are un-vendored (torch_scatter, torch.unique
semantics, scipy CCL): many small random cases against independent stand-ins (torch.unique, scatter_reduce, a float64 loop, scipy)."""
import numpy as np
import torch
from scipy.sparse import csr_matrix
from scipy.sparse.csgraph import connected_components

from oracle import fsf_oracle as O


def test_rank_scatter_and_ccl_against_stand_ins():
    rng = np.random.default_rng(123)
    for _ in range(25):
        n, d = int(rng.integers(1, 300)), int(rng.integers(1, 5))
        rows = rng.integers(0, int(rng.integers(1, 6)), (n, d)).astype(np.int64)
        u, inv, cnt = O.unique_rows(rows)
        tu, tinv, tcnt = torch.unique(torch.from_numpy(rows), dim=0, return_inverse=True, return_counts=True)
        assert np.array_equal(u, tu.numpy()) and np.array_equal(inv, tinv.numpy()) and np.array_equal(cnt, tcnt.numpy())
        c, m = int(rng.integers(1, 9)), int(inv.max()) + 1
        f = (np.round(rng.standard_normal((n, c)) * 2) / 2).astype(np.float32)            # half-integers: many exact ties
        vmax, arg = O.scatter_max(f, inv, m)
        ref = torch.full((m, c), -np.inf).scatter_reduce(0, torch.from_numpy(inv)[:, None].expand(n, c), torch.from_numpy(f), "amax",
                                                         include_self=True).numpy()
        assert np.array_equal(vmax, ref)
        for s in range(m):
            for j in range(c):
                assert arg[s, j] == np.flatnonzero((inv == s) & (f[:, j] == vmax[s, j]))[0]      # lowest row attaining the maximum
        np.testing.assert_allclose(O.scatter_mean(f, inv, m), np.stack([f[inv == s].astype(np.float64).mean(0) for s in range(m)]),
                                   rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(O.scatter_sum(f, inv, m), np.stack([f[inv == s].astype(np.float64).sum(0) for s in range(m)]),
                                   rtol=1e-5, atol=1e-6)
        k, dist = int(rng.integers(1, 150)), float(rng.uniform(0.2, 1.0))
        p = rng.uniform(-3, 3, (k, 3)).astype(np.float32)
        lab = O.connected_components_single_batch(p, dist)
        dxy = np.sqrt(((p[:, None, :2] - p[None, :, :2]) ** 2).sum(-1))
        _, sl = connected_components(csr_matrix(dxy < dist), directed=False)
        first = {}
        assert np.array_equal(lab, np.array([first.setdefault(v, len(first)) for v in sl]))    # same partition, first-seen numbering
