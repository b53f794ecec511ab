"""Drop-in `spconv` / `cumm` module objects backed by libtoda_b200.

    import toda_b200.spconv_compat as sc; sc.install()

registers `spconv`, `spconv.pytorch`, `spconv.pytorch.conv`, `spconv.utils`, `cumm`, `cumm.tensorview`
in sys.modules, so unmodified pcdet files (`import spconv.pytorch as spconv`, `from spconv.utils import
Point2VoxelCPU3d`, `import cumm.tensorview as tv`) bind to the B200 implementation.
"""
import sys
import types

from . import pytorch, utils


def make_modules():
    spconv = types.ModuleType("spconv")
    spconv.__path__ = []
    conv = types.ModuleType("spconv.pytorch.conv")
    conv.SparseConvolution = pytorch.SparseConvolution
    pytorch.conv = conv
    spconv.pytorch = pytorch
    spconv.utils = utils
    cumm = types.ModuleType("cumm")
    cumm.__path__ = []
    tv = types.ModuleType("cumm.tensorview")
    tv.from_numpy = utils.from_numpy
    cumm.tensorview = tv
    return {"spconv": spconv, "spconv.pytorch": pytorch, "spconv.pytorch.conv": conv, "spconv.utils": utils,
            "cumm": cumm, "cumm.tensorview": tv}


def install():
    mods = make_modules()
    sys.modules.update(mods)
    return mods
