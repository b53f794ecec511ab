"""oracle/spconv_oracle.py -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU (pure PyTorch + numpy) restatement of the part of the third-party `spconv`
package that the reference's hot path touches:
  * pcdet/utils/spconv_utils.py L3-6, L19, L29-31 (import, SparseConvolution, replace_feature)
  * pcdet/models/backbones_3d/spconv_backbone.py L12-21, L30, L38-45, L77-117,
    L141-146, L191-232, L254-259 (SubMConv3d / SparseConv3d / SparseSequential /
    SparseModule / SparseConvTensor)
  * pcdet/models/backbones_2d/map_to_bev/height_compression.py L21 (`.dense()`)
`spconv` (+ `cumm`) is un-vendored and un-pinned in the reference (setup.py L48,
docker/Dockerfile L55 `pip install spconv-cu102`, docs/INSTALL.md L9), and is not
installable here, so its published semantics are restated (SURVEY.md Appendix C.2,
C.3).  PARITY UNPINNED by the reference; pinned by dense-equivalence known-answer
tests against torch.nn.functional.conv3d (tests/test_oracle_spconv.py) and by
golden vectors produced through the reference's own spconv_backbone.py /
height_compression.py (tests/golden/).

Everything is built from differentiable torch ops, so torch.autograd supplies the
backward oracle (dgrad, wgrad, bias grad) for free.

May be imported only by tests/, __graft_entry__.smoke() and the CPU-baseline legs
of bench.py.
"""
import sys
import types
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from . import voxelize as _vox


# ----------------------------------------------------------------------------------------------
# geometry
# ----------------------------------------------------------------------------------------------
def _triple(v):
    if isinstance(v, (list, tuple)):
        assert len(v) == 3
        return [int(x) for x in v]
    return [int(v)] * 3


def conv_out_size(in_size, k, s, p, d=1):
    """Appendix C.2: floor((in + 2p - d(k-1) - 1)/s) + 1."""
    return (in_size + 2 * p - d * (k - 1) - 1) // s + 1


def _lin(b, z, y, x, shape):
    d, h, w = shape
    return ((b.astype(np.int64) * d + z) * h + y) * w + x


def _lookup(sorted_keys, order, query):
    """rows of `query` keys in the original (unsorted) array, -1 where absent."""
    pos = np.searchsorted(sorted_keys, query)
    pos_c = np.minimum(pos, len(sorted_keys) - 1)
    hit = (pos < len(sorted_keys)) & (sorted_keys[pos_c] == query)
    return np.where(hit, order[pos_c], -1)


def subm_rulebook(indices, spatial_shape, ksize):
    """Submanifold rulebook: outputs == inputs (same row order); pair (i -> o) under offset k
    iff coord(i) == coord(o) + (k - K//2).  Returns [(in_rows, out_rows)] per kernel offset,
    k = (kz*KH + ky)*KW + kx."""
    idx = np.asarray(indices, dtype=np.int64)
    shape = [int(s) for s in spatial_shape]
    b, z, y, x = idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]
    keys = _lin(b, z, y, x, shape)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    pairs = []
    kz_, ky_, kx_ = ksize
    rows = np.arange(idx.shape[0], dtype=np.int64)
    for kz in range(kz_):
        for ky in range(ky_):
            for kx in range(kx_):
                nz, ny, nx = z + kz - kz_ // 2, y + ky - ky_ // 2, x + kx - kx_ // 2
                ok = (nz >= 0) & (nz < shape[0]) & (ny >= 0) & (ny < shape[1]) & (nx >= 0) & (nx < shape[2])
                q = _lin(b[ok], nz[ok], ny[ok], nx[ok], shape)
                src = _lookup(skeys, order, q) if len(skeys) else np.zeros(0, np.int64)
                out_rows = rows[ok]
                hit = src >= 0
                pairs.append((src[hit].astype(np.int64), out_rows[hit]))
    return pairs


def sparse_rulebook(indices, spatial_shape, ksize, stride, padding):
    """Strided sparse conv rulebook.  Output active set = every o with o*s = p + pad - k for some
    active input p and kernel offset k, 0 <= o < out_shape.  Output rows are emitted in canonical
    (b,z,y,x) lexicographic order (spconv's own order is implementation-defined).
    Returns out_indices (int32 [M,4]), out_shape, [(in_rows, out_rows)] per offset."""
    idx = np.asarray(indices, dtype=np.int64)
    shape = [int(s) for s in spatial_shape]
    out_shape = [conv_out_size(shape[a], ksize[a], stride[a], padding[a]) for a in range(3)]
    b, p = idx[:, 0], [idx[:, 1], idx[:, 2], idx[:, 3]]
    rows = np.arange(idx.shape[0], dtype=np.int64)
    cand = []
    for kz in range(ksize[0]):
        for ky in range(ksize[1]):
            for kx in range(ksize[2]):
                k3 = (kz, ky, kx)
                ok = np.ones(idx.shape[0], bool)
                o = []
                for a in range(3):
                    t = p[a] + padding[a] - k3[a]
                    oa = t // stride[a]
                    ok &= (t >= 0) & (t % stride[a] == 0) & (oa < out_shape[a])
                    o.append(oa)
                key = _lin(b[ok], o[0][ok], o[1][ok], o[2][ok], out_shape)
                cand.append((rows[ok], key))
    all_keys = np.concatenate([c[1] for c in cand]) if cand else np.zeros(0, np.int64)
    ukeys = np.unique(all_keys)  # sorted => canonical (b,z,y,x) order
    d, h, w = out_shape
    ob = ukeys // (d * h * w)
    rem = ukeys % (d * h * w)
    oz = rem // (h * w)
    rem = rem % (h * w)
    out_indices = np.stack([ob, oz, rem // w, rem % w], axis=1).astype(np.int32)
    pairs = []
    for in_rows, key in cand:
        out_rows = np.searchsorted(ukeys, key).astype(np.int64)
        pairs.append((in_rows, out_rows))
    return out_indices, out_shape, pairs


# ----------------------------------------------------------------------------------------------
# spconv.pytorch surface
# ----------------------------------------------------------------------------------------------
class SparseConvTensor:
    def __init__(self, features, indices, spatial_shape, batch_size, grid=None, voxel_num=None,
                 indice_dict=None, benchmark=False):
        self.features = features
        self.indices = indices
        self.spatial_shape = [int(s) for s in spatial_shape]
        self.batch_size = int(batch_size)
        self.indice_dict = indice_dict if indice_dict is not None else {}
        self.grid = grid
        self.voxel_num = voxel_num
        self.benchmark = benchmark

    def replace_feature(self, feature):
        t = SparseConvTensor(feature, self.indices, self.spatial_shape, self.batch_size, self.grid,
                             self.voxel_num, self.indice_dict, self.benchmark)
        return t

    @property
    def spatial_size(self):
        return int(np.prod(self.spatial_shape))

    def dense(self, channels_first=True):
        """Appendix C.3: scatter into zeros (B,D,H,W,C), then permute to (B,C,D,H,W) contiguous."""
        idx = self.indices.long()
        c = self.features.shape[1]
        out = torch.zeros([self.batch_size] + self.spatial_shape + [c], dtype=self.features.dtype,
                          device=self.features.device)
        out = out.index_put((idx[:, 0], idx[:, 1], idx[:, 2], idx[:, 3]), self.features)
        if not channels_first:
            return out
        return out.permute(0, 4, 1, 2, 3).contiguous()


class SparseModule(nn.Module):
    """Marker base class: SparseSequential hands the whole SparseConvTensor to these."""
    pass


class SparseSequential(SparseModule):
    def __init__(self, *args, **kwargs):
        super().__init__()
        if len(args) == 1 and isinstance(args[0], OrderedDict):
            for key, module in args[0].items():
                self.add_module(key, module)
        else:
            for i, module in enumerate(args):
                self.add_module(str(i), module)
        for name, module in kwargs.items():
            self.add_module(name, module)

    def __getitem__(self, idx):
        mods = list(self._modules.values())
        return mods[idx]

    def __len__(self):
        return len(self._modules)

    def add(self, module, name=None):
        self.add_module(name if name is not None else str(len(self._modules)), module)

    def forward(self, x):
        for module in self._modules.values():
            if isinstance(module, SparseModule):
                x = module(x)
            elif isinstance(x, SparseConvTensor):
                if x.indices.shape[0] != 0:
                    x = x.replace_feature(module(x.features))
            else:
                x = module(x)
        return x


class SparseConvolution(SparseModule):
    """weight: (Cout, kz, ky, kx, Cin) -- one of the layouts detector3d_template.py L341-348
    understands; bias: (Cout,)."""

    def __init__(self, ndim, in_channels, out_channels, kernel_size=3, stride=1, padding=0, dilation=1,
                 groups=1, bias=True, subm=False, output_padding=0, transposed=False, inverse=False,
                 indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__()
        assert ndim == 3 and groups == 1 and not transposed and not inverse
        self.ndim = ndim
        self.in_channels = in_channels
        self.out_channels = out_channels
        self.kernel_size = _triple(kernel_size)
        self.stride = _triple(stride)
        self.padding = _triple(padding)
        self.dilation = _triple(dilation)
        assert self.dilation == [1, 1, 1]
        self.subm = subm
        self.indice_key = indice_key
        self.weight = nn.Parameter(torch.empty(out_channels, *self.kernel_size, in_channels))
        if bias:
            self.bias = nn.Parameter(torch.empty(out_channels))
        else:
            self.register_parameter("bias", None)
        self.reset_parameters()

    def reset_parameters(self):
        import math
        nn.init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            fan_in = self.in_channels * int(np.prod(self.kernel_size))
            bound = 1 / math.sqrt(fan_in)
            nn.init.uniform_(self.bias, -bound, bound)

    def rulebook(self, x):
        """-> (out_indices tensor, out_shape, pairs)"""
        key = self.indice_key
        if key is not None and key in x.indice_dict:
            return x.indice_dict[key]
        idx_np = x.indices.detach().cpu().numpy()
        if self.subm:
            assert self.stride == [1, 1, 1]
            pairs = subm_rulebook(idx_np, x.spatial_shape, self.kernel_size)
            rb = (x.indices, list(x.spatial_shape), pairs)
        else:
            oi, oshape, pairs = sparse_rulebook(idx_np, x.spatial_shape, self.kernel_size, self.stride,
                                                self.padding)
            rb = (torch.from_numpy(oi).to(x.indices.device), oshape, pairs)
        if key is not None:
            x.indice_dict[key] = rb
        return rb

    def forward(self, x):
        assert isinstance(x, SparseConvTensor)
        out_indices, out_shape, pairs = self.rulebook(x)
        feats = x.features
        n_out = out_indices.shape[0]
        out = feats.new_zeros((n_out, self.out_channels))
        kz_, ky_, kx_ = self.kernel_size
        k = 0
        for kz in range(kz_):
            for ky in range(ky_):
                for kx in range(kx_):
                    in_rows, out_rows = pairs[k]
                    k += 1
                    if len(in_rows) == 0:
                        continue
                    wk = self.weight[:, kz, ky, kx, :].t()  # (Cin, Cout)
                    src = feats.index_select(0, torch.from_numpy(in_rows)) @ wk
                    out = out.index_add(0, torch.from_numpy(out_rows), src)
        if self.bias is not None:
            out = out + self.bias
        res = SparseConvTensor(out, out_indices, out_shape, x.batch_size, x.grid, x.voxel_num,
                               x.indice_dict, x.benchmark)
        return res


class SubMConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         True, indice_key=indice_key)


class SparseConv3d(SparseConvolution):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 bias=True, indice_key=None, algo=None, fp32_accum=None, name=None):
        super().__init__(3, in_channels, out_channels, kernel_size, stride, padding, dilation, groups, bias,
                         False, indice_key=indice_key)


class SparseInverseConv3d(SparseModule):
    """Referenced only by the unused 'inverseconv' branch of post_act_block (spconv_backbone.py L16-17)."""

    def __init__(self, *a, **kw):
        super().__init__()
        raise NotImplementedError("SparseInverseConv3d is not on the hot path (SURVEY.md section 8b)")


def rulebook_pairs_canonical(indices_in, indices_out, pairs, in_shape, out_shape):
    """Canonical form of a rulebook for parity checks: int64 array of rows
    (k, out_key, in_key) sorted lexicographically, keys being linear (b,z,y,x) cell ids."""
    ii = np.asarray(indices_in, dtype=np.int64)
    oi = np.asarray(indices_out, dtype=np.int64)
    kin = _lin(ii[:, 0], ii[:, 1], ii[:, 2], ii[:, 3], in_shape)
    kout = _lin(oi[:, 0], oi[:, 1], oi[:, 2], oi[:, 3], out_shape)
    rows = []
    for k, (a, b) in enumerate(pairs):
        if len(a):
            rows.append(np.stack([np.full(len(a), k, np.int64), kout[b], kin[a]], axis=1))
    if not rows:
        return np.zeros((0, 3), np.int64)
    allr = np.concatenate(rows)
    order = np.lexsort((allr[:, 2], allr[:, 1], allr[:, 0]))
    return allr[order]


# ----------------------------------------------------------------------------------------------
# install as `spconv` / `cumm` so that the reference's own files import against the oracle
# ----------------------------------------------------------------------------------------------
def make_modules():
    """Module objects shaped like the names the reference imports."""
    this = sys.modules[__name__]
    spconv = types.ModuleType("spconv")
    spconv.__path__ = []
    pt = types.ModuleType("spconv.pytorch")
    pt.__path__ = []
    conv = types.ModuleType("spconv.pytorch.conv")
    for name in ["SparseConvTensor", "SparseModule", "SparseSequential", "SparseConvolution", "SubMConv3d",
                 "SparseConv3d", "SparseInverseConv3d"]:
        setattr(pt, name, getattr(this, name))
    conv.SparseConvolution = SparseConvolution
    pt.conv = conv
    utils = types.ModuleType("spconv.utils")
    utils.Point2VoxelCPU3d = _vox.Point2VoxelCPU3d
    spconv.pytorch = pt
    spconv.utils = utils
    cumm = types.ModuleType("cumm")
    cumm.__path__ = []
    tv = types.ModuleType("cumm.tensorview")
    tv.from_numpy = _vox.from_numpy
    cumm.tensorview = tv
    return {"spconv": spconv, "spconv.pytorch": pt, "spconv.pytorch.conv": conv, "spconv.utils": utils,
            "cumm": cumm, "cumm.tensorview": tv}


def install():
    mods = make_modules()
    sys.modules.update(mods)
    return mods
