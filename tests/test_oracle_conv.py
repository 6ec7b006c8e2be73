"""The oracle's sparse-convolution rulebook + gather-GEMM against torch's dense conv3d /
conv_transpose3d on a densified grid (the executable stand-in for the absent spconv)."""
import numpy as np
import torch

from oracle import fsf_oracle as O


def _scene(seed=0, shape=(2, 6, 10, 12), density=0.25):
    rng = np.random.default_rng(seed)
    dense = rng.random(shape) < density
    return rng, shape, np.argwhere(dense).astype(np.int32)


def _densify(coors, feat, shape):
    x = torch.zeros(shape[0], feat.shape[1], *shape[1:])
    x[coors[:, 0], :, coors[:, 1], coors[:, 2], coors[:, 3]] = torch.from_numpy(feat)
    return x


def _at(y, coors):
    return y[coors[:, 0], :, coors[:, 1], coors[:, 2], coors[:, 3]].numpy()


def test_subm_conv_matches_dense():
    rng, shape, coors = _scene()
    cin, cout = 3, 4
    feat = rng.standard_normal((len(coors), cin)).astype(np.float32)
    w = rng.standard_normal((27, cout, cin)).astype(np.float32)
    nbr = O.conv_rulebook(coors, coors, shape, (3, 3, 3), (1, 1, 1), (1, 1, 1))
    assert np.array_equal(nbr[13], np.arange(len(coors)))  # centre offset = identity
    out = O.gather_gemm(feat, w, nbr)
    wt = torch.from_numpy(w).reshape(3, 3, 3, cout, cin).permute(3, 4, 0, 1, 2)
    y = torch.nn.functional.conv3d(_densify(coors, feat, shape), wt, padding=1)
    np.testing.assert_allclose(out, _at(y, coors), rtol=1e-5, atol=1e-5)


def test_strided_and_inverse_conv_match_dense():
    rng, shape, coors = _scene(1)
    cin, cout = 3, 5
    feat = rng.standard_normal((len(coors), cin)).astype(np.float32)
    w = rng.standard_normal((27, cout, cin)).astype(np.float32)
    for pad in [(1, 1, 1), (0, 1, 1)]:
        oshape = (shape[0],) + tuple((shape[1 + a] + 2 * pad[a] - 3) // 2 + 1 for a in range(3))
        oc = O.conv_out_coors(coors, oshape, (3, 3, 3), (2, 2, 2), pad)
        x = _densify(coors, feat, shape)
        wt = torch.from_numpy(w).reshape(3, 3, 3, cout, cin).permute(3, 4, 0, 1, 2)
        y = torch.nn.functional.conv3d(x, wt, padding=pad, stride=2)
        act = torch.nn.functional.conv3d((x.abs().sum(1, keepdim=True) > 0).float(), torch.ones(1, 1, 3, 3, 3), padding=pad,
                                         stride=2)[:, 0] > 0
        assert np.array_equal(np.argwhere(act.numpy()), oc)
        nbr = O.conv_rulebook(oc, coors, shape, (3, 3, 3), (2, 2, 2), pad)
        np.testing.assert_allclose(O.gather_gemm(feat, w, nbr), _at(y, oc), rtol=1e-5, atol=1e-5)
        # inverse conv back onto the fine set == conv_transpose3d sampled at the fine sites
        f2 = rng.standard_normal((len(oc), cout)).astype(np.float32)
        w3 = rng.standard_normal((27, cin, cout)).astype(np.float32)
        nbr3 = O.conv_rulebook(coors, oc, oshape, (3, 3, 3), (2, 2, 2), pad, transposed=True)
        wt3 = torch.from_numpy(w3).reshape(3, 3, 3, cin, cout).permute(4, 3, 0, 1, 2)
        opad = tuple(shape[1 + a] - ((oshape[1 + a] - 1) * 2 - 2 * pad[a] + 3) for a in range(3))
        y3 = torch.nn.functional.conv_transpose3d(_densify(oc, f2, oshape), wt3, stride=2, padding=pad, output_padding=opad)
        np.testing.assert_allclose(O.gather_gemm(f2, w3, nbr3), _at(y3, coors), rtol=1e-5, atol=1e-5)
        # forward and inverse rulebooks hold the same pairs
        fwd = {(int(nbr[k, o]), o, k) for k in range(27) for o in range(len(oc)) if nbr[k, o] >= 0}
        inv = {(i, int(nbr3[k, i]), k) for k in range(27) for i in range(len(coors)) if nbr3[k, i] >= 0}
        assert fwd == inv
