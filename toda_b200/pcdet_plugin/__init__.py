"""pcdet plugin surface: the cfg-driven module NAMEs of the hot path, backed by libtoda_b200.

Registries mirror pcdet's (`vfe.__all__` pcdet/models/backbones_3d/vfe/__init__.py L8-15,
`backbones_3d.__all__` backbones_3d/__init__.py L5-11, `map_to_bev.__all__` map_to_bev/__init__.py L5-9);
`register(pcdet_models)` overwrites those entries in an imported pcdet so that `build_network` picks these
classes up with unchanged YAML configs.
"""
import torch
import torch.nn as nn

from .. import ops
from ..spconv_compat import pytorch as _sp
from .backbones import make_backbones

_classes = make_backbones(_sp, _sp.bn_act_tensor, _sp.conv_bn_act_tensor, _sp.res_block_tensor)
VoxelBackBone8x = _classes["VoxelBackBone8x"]
VoxelResBackBone8x = _classes["VoxelResBackBone8x"]
SparseBasicBlock = _classes["SparseBasicBlock"]


class MeanVFE(nn.Module):
    """pcdet/models/backbones_3d/vfe/mean_vfe.py; constructor as detector3d_template.py L56-63 calls it.

    With the reference's `transform_points_to_voxels` processor the batch already holds voxels / voxel_coords /
    voxel_num_points and only the mean is computed.  With `transform_points_to_voxels_placeholder`
    (data_processor.py L105-113) the batch holds `points` [b,x,y,z,...] on the GPU and the hard voxelization
    runs here (K1), which is where the work belongs once it is a GPU kernel; the cap parameters then come
    from model_cfg (MAX_POINTS_PER_VOXEL, MAX_NUMBER_OF_VOXELS) since pcdet does not hand the data config to
    the VFE.
    """

    def __init__(self, model_cfg, num_point_features, voxel_size=None, point_cloud_range=None, grid_size=None, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_point_features = num_point_features
        self.voxel_size = voxel_size
        self.point_cloud_range = point_cloud_range
        self.grid_size = grid_size

    def get_output_feature_dim(self):
        return self.num_point_features

    def _cfg(self, key, default=None):
        cfg = self.model_cfg
        val = cfg.get(key, default) if hasattr(cfg, "get") else getattr(cfg, key, default)
        if val is None:
            raise KeyError(f"MeanVFE needs model_cfg.{key} to voxelize on the GPU")
        return val

    def voxelize(self, batch_dict):
        points = batch_dict["points"]                       # (N, 1+F): [b, x, y, z, ...]
        batch_size = int(batch_dict["batch_size"])
        offsets = batch_dict.get("point_frame_offsets")
        if offsets is None:
            # frames are stored back to back by collate_batch (dataset.py L173-178)
            bidx = points[:, 0].contiguous()
            bounds = torch.arange(batch_size + 1, device=points.device, dtype=torch.float32)
            offsets = torch.searchsorted(bidx, bounds).int()
        mode = "train" if self.training else "test"
        max_voxels = self._cfg("MAX_NUMBER_OF_VOXELS")
        max_voxels = max_voxels[mode] if hasattr(max_voxels, "__getitem__") and not isinstance(max_voxels, int) else max_voxels
        voxels, coords, num, _ = ops.voxelize(
            points.contiguous(), offsets, self.point_cloud_range, self.voxel_size, int(self._cfg("MAX_POINTS_PER_VOXEL")),
            int(max_voxels), num_features=self.num_point_features, xyz_col=1, feat_col=1, order=ops.ORDER_CANONICAL,
            grid=self.grid_size)
        batch_dict["voxels"], batch_dict["voxel_coords"], batch_dict["voxel_num_points"] = voxels, coords, num
        batch_dict["voxel_coords_canonical"] = True
        return batch_dict

    def forward(self, batch_dict, **kwargs):
        if "voxels" not in batch_dict:
            batch_dict = self.voxelize(batch_dict)
        batch_dict["voxel_features"] = ops.mean_vfe(batch_dict["voxels"], batch_dict["voxel_num_points"])
        return batch_dict


class HeightCompression(nn.Module):
    """pcdet/models/backbones_2d/map_to_bev/height_compression.py: one scatter kernel writes (B, C*D, H, W)."""

    def __init__(self, model_cfg, **kwargs):
        super().__init__()
        self.model_cfg = model_cfg
        self.num_bev_features = self.model_cfg.NUM_BEV_FEATURES

    def forward(self, batch_dict):
        t = batch_dict["encoded_spconv_tensor"]
        d, h, w = t.spatial_shape
        idx = t.indices if t.indices.dtype == torch.int32 else t.indices.int()
        batch_dict["spatial_features"] = ops.bev_scatter(t.features, idx.contiguous(), t.batch_size, d, h, w)
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        return batch_dict


VFE = {"MeanVFE": MeanVFE}
BACKBONES_3D = {"VoxelBackBone8x": VoxelBackBone8x, "VoxelResBackBone8x": VoxelResBackBone8x}
MAP_TO_BEV = {"HeightCompression": HeightCompression}


def register(pcdet_models):
    """Overwrite the hot-path entries of an imported `pcdet.models` package's registries."""
    pcdet_models.backbones_3d.vfe.__all__.update(VFE)
    pcdet_models.backbones_3d.__all__.update(BACKBONES_3D)
    pcdet_models.backbones_2d.map_to_bev.__all__.update(MAP_TO_BEV)
