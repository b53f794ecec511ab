"""Known-answer tests that pin the CPU spconv oracle (oracle/spconv_oracle.py) independently of
any recollection of spconv internals: dense equivalence against
torch.nn.functional.conv3d on the densified tensor, for values, active sets and gradients."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import spconv_oracle as S

torch.manual_seed(0)


def random_sparse(seed, batch, shape, n, c, positive=False):
    rng = np.random.default_rng(seed)
    d, h, w = shape
    cells = rng.choice(batch * d * h * w, size=n, replace=False)
    b = cells // (d * h * w)
    r = cells % (d * h * w)
    idx = np.stack([b, r // (h * w), (r % (h * w)) // w, r % w], axis=1).astype(np.int32)
    feats = rng.standard_normal((n, c)).astype(np.float32)
    if positive:
        feats = np.abs(feats) + 0.1
    return torch.from_numpy(feats), torch.from_numpy(idx)


def dense_of(feats, idx, batch, shape):
    return S.SparseConvTensor(feats, idx, shape, batch).dense()


def dense_weight(conv):
    # (Cout,kz,ky,kx,Cin) -> conv3d's (Cout,Cin,kz,ky,kx)
    return conv.weight.permute(0, 4, 1, 2, 3).contiguous()


@pytest.mark.parametrize("cin,cout,bias", [(5, 16, False), (16, 16, True), (8, 24, True)])
def test_subm_dense_equivalence_fwd_bwd(cin, cout, bias):
    batch, shape = 2, [7, 9, 11]
    feats, idx = random_sparse(1, batch, shape, 300, cin)
    conv = S.SubMConv3d(cin, cout, 3, padding=1, bias=bias, indice_key="k")
    f1 = feats.clone().requires_grad_(True)
    out = conv(S.SparseConvTensor(f1, idx, shape, batch))
    assert torch.equal(out.indices, idx) and out.spatial_shape == shape      # outputs == inputs, same order
    f2 = feats.clone().requires_grad_(True)
    dense_out = F.conv3d(dense_of(f2, idx, batch, shape), dense_weight(conv), conv.bias, stride=1, padding=1)
    li = idx.long()
    ref = dense_out[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]]
    torch.testing.assert_close(out.features, ref, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(ref)
    gw1 = torch.autograd.grad((out.features * g).sum(), [f1, conv.weight] + ([conv.bias] if bias else []))
    gw2 = torch.autograd.grad((ref * g).sum(), [f2, conv.weight] + ([conv.bias] if bias else []))
    for a, b in zip(gw1, gw2):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("k,s,p", [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                                   ((3, 1, 1), (2, 1, 1), (0, 0, 0)), ((3, 3, 3), (1, 1, 1), (0, 0, 0))])
def test_sparse_conv_dense_equivalence(k, s, p):
    batch, shape, cin, cout = 2, [9, 10, 12], 6, 10
    feats, idx = random_sparse(2, batch, shape, 250, cin, positive=True)
    conv = S.SparseConv3d(cin, cout, k, stride=s, padding=p, bias=False, indice_key="d")
    with torch.no_grad():
        conv.weight.abs_().add_(0.05)          # positive weights x positive inputs: active <=> non-zero
    f1 = feats.clone().requires_grad_(True)
    out = conv(S.SparseConvTensor(f1, idx, shape, batch))
    f2 = feats.clone().requires_grad_(True)
    dense_out = F.conv3d(dense_of(f2, idx, batch, shape), dense_weight(conv), None, stride=s, padding=p)
    assert list(dense_out.shape[2:]) == out.spatial_shape
    # active set: exactly the non-zero sites of the dense result, in canonical (b,z,y,x) order
    nz = (dense_out.abs().sum(1) > 0).nonzero()
    assert torch.equal(nz.int(), out.indices.int())
    li = out.indices.long()
    ref = dense_out[li[:, 0], :, li[:, 1], li[:, 2], li[:, 3]]
    torch.testing.assert_close(out.features, ref, rtol=1e-4, atol=1e-5)
    g = torch.randn_like(ref)
    ga = torch.autograd.grad((out.features * g).sum(), [f1, conv.weight])
    gb = torch.autograd.grad((ref * g).sum(), [f2, conv.weight])
    for a, b in zip(ga, gb):
        torch.testing.assert_close(a, b, rtol=1e-4, atol=1e-4)


def test_shape_kats_from_survey():
    # [41,1440,1440]->[21,720,720]->[11,360,360]->[5,180,180]->[2,180,180]
    sh = [41, 1440, 1440]
    for k, s, p, want in [((3, 3, 3), (2, 2, 2), (1, 1, 1), [21, 720, 720]),
                          ((3, 3, 3), (2, 2, 2), (1, 1, 1), [11, 360, 360]),
                          ((3, 3, 3), (2, 2, 2), (0, 1, 1), [5, 180, 180]),
                          ((3, 1, 1), (2, 1, 1), (0, 0, 0), [2, 180, 180])]:
        sh = [S.conv_out_size(sh[a], k[a], s[a], p[a]) for a in range(3)]
        assert sh == want


def test_indice_key_reuse_and_sequential():
    batch, shape = 1, [5, 6, 7]
    feats, idx = random_sparse(3, batch, shape, 60, 4)
    seq = S.SparseSequential(S.SubMConv3d(4, 8, 3, padding=1, bias=False, indice_key="a"),
                             torch.nn.BatchNorm1d(8), torch.nn.ReLU(),
                             S.SubMConv3d(8, 8, 3, padding=1, bias=True, indice_key="a"))
    x = S.SparseConvTensor(feats, idx, shape, batch)
    y = seq(x)
    assert "a" in x.indice_dict and y.features.shape == (60, 8) and len(seq) == 4
    assert y.indice_dict is x.indice_dict


def test_dense_layout_matches_height_compression_view():
    batch, shape = 2, [2, 4, 5]
    feats, idx = random_sparse(4, batch, shape, 30, 3)
    d = S.SparseConvTensor(feats, idx, shape, batch).dense()
    assert d.shape == (2, 3, 2, 4, 5) and d.is_contiguous()
    bev = d.view(2, 3 * 2, 4, 5)          # height_compression.py L22-23: channel = c*D + z
    for n in range(30):
        b, z, y, x = idx[n].tolist()
        for c in range(3):
            assert bev[b, c * 2 + z, y, x] == feats[n, c]
    assert (bev != 0).sum() == (feats != 0).sum()
