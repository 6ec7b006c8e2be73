"""GPU parity of the sparse-convolution rulebook (bit-exact) and SubM / strided / inverse
convolutions through the tcgen05 gather-GEMM (1e-4)."""
import numpy as np
import pytest
import torch

from fullysparsefusion_b200 import ops, synth
from oracle import fsf_oracle as O

pytestmark = pytest.mark.gpu


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


def _voxels(n, seed, batches=1, sweeps=1):
    pts = synth.ring_points(n, sweeps=sweeps, seed=seed)
    c = O.voxelize(pts, synth.NUSC_VOXEL, synth.NUSC_RANGE, floor_mode=0).astype(np.int64)
    b = np.random.default_rng(seed).integers(0, batches, (n, 1))
    return O.unique_rows(np.concatenate([b, c], 1))[0].astype(np.int32)


@pytest.mark.parametrize("n,batches", [(50, 1), (5000, 2), (34000, 1), (300000, 1)])
def test_subm_rulebook(cuda, n, batches):
    coors = _voxels(n, 3, batches, sweeps=10 if n > 100000 else 1)
    shape = (batches, 40, 512, 512)
    dc = T(coors, cuda)
    _, inv, _, index = ops.unique_rows(dc, lo=[0, 0, 0, 0], ext=list(shape), return_index=True)
    assert torch.equal(inv.cpu(), torch.arange(len(coors)))
    nbr = ops.conv_rulebook(dc, index, 3, 1, 1).cpu().numpy()
    want = O.conv_rulebook(coors, coors, shape, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(nbr, want)
    assert np.array_equal(nbr[13], np.arange(len(coors)))


@pytest.mark.parametrize("pad", [(1, 1, 1), (0, 1, 1)])
@pytest.mark.parametrize("n", [3000, 30000])
def test_strided_and_inverse_rulebook(cuda, n, pad):
    coors = _voxels(n, 7, 2)
    shape = (2, 40, 512, 512)
    oshape = (2,) + tuple((shape[1 + a] + 2 * pad[a] - 3) // 2 + 1 for a in range(3))
    dc = T(coors, cuda)
    index = ops.unique_rows(dc, lo=[0, 0, 0, 0], ext=list(shape), return_index=True)[3]
    oc, oindex = ops.conv_out_index(dc, oshape, 3, 2, pad)
    want_oc = O.conv_out_coors(coors, oshape, (3, 3, 3), (2, 2, 2), pad)
    assert np.array_equal(oc.cpu().numpy(), want_oc)
    nbr = ops.conv_rulebook(oc, index, 3, 2, pad).cpu().numpy()
    assert np.array_equal(nbr, O.conv_rulebook(want_oc, coors, shape, (3, 3, 3), (2, 2, 2), pad))
    nbr_t = ops.conv_rulebook(dc, oindex, 3, 2, pad, transposed=True).cpu().numpy()
    assert np.array_equal(nbr_t, O.conv_rulebook(coors, want_oc, oshape, (3, 3, 3), (2, 2, 2), pad, transposed=True))


def test_conv_layers_vs_oracle(cuda):
    """SubM conv → BN(affine) → ReLU, strided conv, inverse conv with residual: values vs oracle."""
    rng = np.random.default_rng(0)
    coors = _voxels(6000, 11)
    shape, oshape = (1, 40, 512, 512), (1, 20, 256, 256)
    dc = T(coors, cuda)
    index = ops.unique_rows(dc, lo=[0, 0, 0, 0], ext=list(shape), return_index=True)[3]
    m, cin, cout = len(coors), 16, 32
    feat = rng.standard_normal((m, cin)).astype(np.float32)
    w = (rng.standard_normal((27, cout, cin)) / np.sqrt(cin * 9)).astype(np.float32)
    scale, shift = rng.uniform(0.5, 1.5, cout).astype(np.float32), rng.standard_normal(cout).astype(np.float32)
    nbr = ops.conv_rulebook(dc, index, 3, 1, 1)
    got = ops.gather_gemm(T(feat, cuda), ops.gemm_prepack(T(w, cuda)), nbr=nbr, norm="affine", norm_w=T(scale, cuda),
                          norm_b=T(shift, cuda), act="relu")
    want = O.gather_gemm(feat, w, nbr.cpu().numpy(), norm="affine", norm_w=scale, norm_b=shift, act="relu")
    np.testing.assert_allclose(got.cpu().numpy(), want, rtol=1e-4, atol=2e-5)
    # down
    oc, oindex = ops.conv_out_index(dc, oshape, 3, 2, 1)
    nbr_d = ops.conv_rulebook(oc, index, 3, 2, 1)
    w2 = (rng.standard_normal((27, 64, cout)) / np.sqrt(cout * 9)).astype(np.float32)
    down = ops.gather_gemm(got, ops.gemm_prepack(T(w2, cuda)), nbr=nbr_d, act="relu")
    want_d = O.gather_gemm(want, w2, nbr_d.cpu().numpy(), act="relu")
    np.testing.assert_allclose(down.cpu().numpy(), want_d, rtol=1e-4, atol=2e-5)
    # up (inverse) + residual
    nbr_u = ops.conv_rulebook(dc, oindex, 3, 2, 1, transposed=True)
    w3 = (rng.standard_normal((27, cout, 64)) / np.sqrt(64 * 4)).astype(np.float32)
    up = ops.gather_gemm(down, ops.gemm_prepack(T(w3, cuda)), nbr=nbr_u, residual=got, act="relu")
    want_u = O.gather_gemm(want_d, w3, nbr_u.cpu().numpy(), residual=want, act="relu")
    np.testing.assert_allclose(up.cpu().numpy(), want_u, rtol=1e-4, atol=5e-5)
