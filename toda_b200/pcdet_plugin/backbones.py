"""VoxelBackBone8x / VoxelResBackBone8x for the pcdet plugin surface.

Same cfg NAMEs, constructor signature `(model_cfg, input_channels, grid_size, **kwargs)`, attributes
(`num_point_features`, `backbone_channels`, `sparse_shape`), batch_dict keys and state_dict parameter names as
pcdet/models/backbones_3d/spconv_backbone.py (VoxelBackBone8x L69-180, VoxelResBackBone8x L183-293,
SparseBasicBlock L30-66, post_act_block L8-27), so existing configs and checkpoints keep working.
The topology is declared as data (STAGES_*) and instantiated against a sparse-conv provider: the B200
implementation (toda_b200.spconv_compat.pytorch) in production, the CPU oracle in the parity tests.
BatchNorm1d(eps=1e-3, momentum=0.01) + ReLU (+ residual) run as one fused pass (`bn_act`).
"""
from functools import partial

import numpy as np
import torch.nn as nn

# (kind, out_channels, kernel, stride, padding, indice_key)
#   'subm'  : SubMConv3d + BN + ReLU                (post_act_block, conv_type='subm')
#   'down'  : SparseConv3d + BN + ReLU              (post_act_block, conv_type='spconv')
#   'res'   : SparseBasicBlock (two SubMConv3d with bias, BN, residual)
STAGES_VOXEL = dict(
    input=(16, "subm1"),
    conv1=[("subm", 16, 3, 1, 1, "subm1")],
    conv2=[("down", 32, 3, 2, 1, "spconv2"), ("subm", 32, 3, 1, 1, "subm2"), ("subm", 32, 3, 1, 1, "subm2")],
    conv3=[("down", 64, 3, 2, 1, "spconv3"), ("subm", 64, 3, 1, 1, "subm3"), ("subm", 64, 3, 1, 1, "subm3")],
    conv4=[("down", 64, 3, 2, (0, 1, 1), "spconv4"), ("subm", 64, 3, 1, 1, "subm4"), ("subm", 64, 3, 1, 1, "subm4")],
    out=(128, (3, 1, 1), (2, 1, 1), "spconv_down2"),
    channels=dict(x_conv1=16, x_conv2=32, x_conv3=64, x_conv4=64),
)
STAGES_RES = dict(
    input=(16, "subm1"),
    conv1=[("res", 16, 3, 1, 1, "res1"), ("res", 16, 3, 1, 1, "res1")],
    conv2=[("down", 32, 3, 2, 1, "spconv2"), ("res", 32, 3, 1, 1, "res2"), ("res", 32, 3, 1, 1, "res2")],
    conv3=[("down", 64, 3, 2, 1, "spconv3"), ("res", 64, 3, 1, 1, "res3"), ("res", 64, 3, 1, 1, "res3")],
    conv4=[("down", 128, 3, 2, (0, 1, 1), "spconv4"), ("res", 128, 3, 1, 1, "res4"), ("res", 128, 3, 1, 1, "res4")],
    out=(128, (3, 1, 1), (2, 1, 1), "spconv_down2"),
    channels=dict(x_conv1=16, x_conv2=32, x_conv3=64, x_conv4=128),
)


def make_backbones(sp, bn_act, conv_bn_act=None, res_block=None):
    """sp: module with SparseConvTensor, SparseModule, SparseSequential, SubMConv3d, SparseConv3d.
    bn_act(sparse_tensor, bn_module, residual_features, relu) -> sparse_tensor with the new features.
    conv_bn_act(sparse_tensor, conv_module, bn_module, residual_features, relu) (optional): the same result as
    bn_act(conv(x), ...) from one fused call."""
    if conv_bn_act is None:
        def conv_bn_act(x, conv, bn, residual, relu):
            return bn_act(conv(x), bn, residual, relu)
    norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)

    class ConvBNReLU(sp.SparseSequential):
        """conv + BN + ReLU with state_dict keys '0.*' (conv) and '1.*' (BN), as post_act_block produces."""

        def __init__(self, conv, channels):
            super().__init__(conv, norm_fn(channels), nn.ReLU())

        def forward(self, x):
            return conv_bn_act(x, self[0], self[1], None, True)

    class SparseBasicBlock(sp.SparseModule):
        expansion = 1

        def __init__(self, inplanes, planes, indice_key=None):
            super().__init__()
            self.conv1 = sp.SubMConv3d(inplanes, planes, kernel_size=3, stride=1, padding=1, bias=True, indice_key=indice_key)
            self.bn1 = norm_fn(planes)
            self.relu = nn.ReLU()
            self.conv2 = sp.SubMConv3d(planes, planes, kernel_size=3, stride=1, padding=1, bias=True, indice_key=indice_key)
            self.bn2 = norm_fn(planes)

        def forward(self, x):
            if hasattr(x, "canonical"):
                x = x.canonical()      # the residual rows must line up with the conv outputs
            if res_block is not None and self.conv1.indice_key is not None and self.conv1.indice_key == self.conv2.indice_key:
                return res_block(x, self.conv1, self.bn1, self.conv2, self.bn2)     # the whole block as one autograd node
            out = conv_bn_act(x, self.conv1, self.bn1, None, True)
            return conv_bn_act(out, self.conv2, self.bn2, x.features, True)

    def build_stage(cin, spec):
        mods = []
        for kind, cout, k, s, p, key in spec:
            if kind == "subm":
                mods.append(ConvBNReLU(sp.SubMConv3d(cin, cout, k, padding=p, bias=False, indice_key=key), cout))
            elif kind == "down":
                mods.append(ConvBNReLU(sp.SparseConv3d(cin, cout, k, stride=s, padding=p, bias=False, indice_key=key), cout))
            elif kind == "res":
                assert cin == cout
                mods.append(SparseBasicBlock(cin, cout, indice_key=key))
            else:
                raise ValueError(kind)
            cin = cout
        return sp.SparseSequential(*mods), cin

    class _Backbone8x(nn.Module):
        STAGES = None

        def __init__(self, model_cfg, input_channels, grid_size, **kwargs):
            super().__init__()
            self.model_cfg = model_cfg
            st = self.STAGES
            self.sparse_shape = np.asarray(grid_size)[::-1] + [1, 0, 0]
            c0, key0 = st["input"]
            self.conv_input = ConvBNReLU(sp.SubMConv3d(input_channels, c0, 3, padding=1, bias=False, indice_key=key0), c0)
            c = c0
            self.conv1, c = build_stage(c, st["conv1"])
            self.conv2, c = build_stage(c, st["conv2"])
            self.conv3, c = build_stage(c, st["conv3"])
            self.conv4, c = build_stage(c, st["conv4"])
            cout, k, s, key = st["out"]
            last_pad = model_cfg.get("last_pad", 0) if hasattr(model_cfg, "get") else 0
            self.conv_out = ConvBNReLU(sp.SparseConv3d(c, cout, k, stride=s, padding=last_pad, bias=False, indice_key=key), cout)
            self.num_point_features = cout
            self.backbone_channels = dict(st["channels"])

        def plan_geometry(self, voxel_coords, batch_size, canonical=False):
            """Coordinate-only part of forward(): the index and every rulebook, as a plan that forward() picks up from
            batch_dict['spconv_geometry'] (toda_b200.pipeline builds it one batch ahead on a side stream)."""
            convs = [m for m in self.modules() if isinstance(m, sp.SparseConvolution)]     # registration = execution order
            return sp.plan_geometry(convs, voxel_coords, self.sparse_shape, batch_size, assume_canonical=canonical)

        def forward(self, batch_dict):
            voxel_features, voxel_coords = batch_dict["voxel_features"], batch_dict["voxel_coords"]
            plan = batch_dict.get("spconv_geometry")
            if plan is not None:
                if plan.coords.shape[0] != voxel_features.shape[0]:
                    raise ValueError("spconv_geometry was planned for a different batch")
                x = plan.sparse_tensor(voxel_features)
            else:
                x = sp.SparseConvTensor(features=voxel_features, indices=voxel_coords.int(), spatial_shape=self.sparse_shape,
                                        batch_size=batch_dict["batch_size"])
                if batch_dict.get("voxel_coords_canonical", False) and hasattr(x, "canonical"):
                    x = x.canonical(assume_canonical=True)
            x = self.conv_input(x)
            x_conv1 = self.conv1(x)
            x_conv2 = self.conv2(x_conv1)
            x_conv3 = self.conv3(x_conv2)
            x_conv4 = self.conv4(x_conv3)
            out = self.conv_out(x_conv4)
            batch_dict.update({
                "encoded_spconv_tensor": out,
                "encoded_spconv_tensor_stride": 8,
                "multi_scale_3d_features": {"x_conv1": x_conv1, "x_conv2": x_conv2, "x_conv3": x_conv3, "x_conv4": x_conv4},
                "multi_scale_3d_strides": {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8},
            })
            return batch_dict

    class VoxelBackBone8x(_Backbone8x):
        STAGES = STAGES_VOXEL

    class VoxelResBackBone8x(_Backbone8x):
        STAGES = STAGES_RES

    return dict(VoxelBackBone8x=VoxelBackBone8x, VoxelResBackBone8x=VoxelResBackBone8x, SparseBasicBlock=SparseBasicBlock,
                ConvBNReLU=ConvBNReLU)
