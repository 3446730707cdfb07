"""u2mkd_b200 — B200-native (sm_100a) LiDAR point-voxel backbone hot path of U2MKD.

    import u2mkd_b200
    u2mkd_b200.install_as_torchsparse()      # `import torchsparse` now resolves here
    from core.models.semantickitti.spvcnn import SPVCNN   # reference model, unchanged
"""
import importlib
import sys

__version__ = "0.1.0"


def install_as_torchsparse() -> None:
    """Register u2mkd_b200.torchsparse (and its submodules) as `torchsparse` in sys.modules."""
    pkg = importlib.import_module(__name__ + ".torchsparse")
    prefix = pkg.__name__
    for name, mod in list(sys.modules.items()):
        if name == prefix or name.startswith(prefix + "."):
            sys.modules["torchsparse" + name[len(prefix):]] = mod


def install_as_sptr() -> None:
    """Register u2mkd_b200.sptr as `sptr` (third_party/SparseTransformer) in sys.modules: the SphereFormer blocks of
    core/models/sphereformer/spherical_transformer.py then import it unchanged."""
    pkg = importlib.import_module(__name__ + ".sptr")
    prefix = pkg.__name__
    for name, mod in list(sys.modules.items()):
        if name == prefix or name.startswith(prefix + "."):
            sys.modules["sptr" + name[len(prefix):]] = mod


def install_reference_shims() -> list:
    """timm.models.layers / torch_scatter / torchpack.utils.config stand-ins + the third_party.SparseTransformer.sptr import
    path, for the reference's SphereFormer model files (u2mkd_b200/shims)."""
    from .shims import install_reference_shims as _install
    return _install()


def set_math(mode: str) -> None:
    from . import ops
    ops.set_math(mode)
