import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from tests import parity_utils as PU
from oracle import spconv_oracle as S
from toda_b200.spconv_compat import pytorch as G
from toda_b200 import ops
for shape in ([6, 12, 12], [5, 12, 12], [8, 12, 12]):
    for (k, s, p) in (((3, 1, 1), (2, 1, 1), 0), ((3, 3, 3), (2, 2, 2), (0, 1, 1))):
        torch.manual_seed(0)
        a = S.SparseConv3d(16, 16, k, stride=s, padding=p, bias=False, indice_key="a")
        b = G.SparseConv3d(16, 16, k, stride=s, padding=p, bias=False, indice_key="a")
        b.load_state_dict(a.state_dict()); b = b.cuda()
        feats, idx = PU.random_sparse(1, 2, shape, 600, 16)
        fa = torch.from_numpy(feats).requires_grad_(True); fb = torch.from_numpy(feats).cuda().requires_grad_(True)
        ya = a(S.SparseConvTensor(fa, torch.from_numpy(idx), shape, 2))
        yb = b(G.SparseConvTensor(fb, torch.from_numpy(idx).cuda(), shape, 2))
        g = torch.randn(ya.features.shape, generator=torch.Generator().manual_seed(3))
        ya.features.backward(g); yb.features.backward(g.cuda())
        d = (fb.grad.cpu() - fa.grad).abs()
        zbad = np.unique(idx[(d.max(1).values > 1e-4).numpy(), 1])
        print(shape, k, s, p, "fwd err %.2e dgrad err %.2e  bad z planes %s  out shape %s" % (float((yb.features.cpu() - ya.features).abs().max()), float(d.max()), zbad, yb.spatial_shape))
