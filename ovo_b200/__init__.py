"""ovo_b200 — B200-native (sm_100a) hot path of tberriel/OVO behind the reference's own entity API.

    from ovo_b200 import OVO, CLIPGenerator, MaskGenerator, Instance3D      # drop-in for ovo/entities/*

Everything that computes runs in libovo_b200.so (include/ovo_b200.h); there is no CPU fallback."""
from .instance3d import Instance3D  # noqa: F401

__all__ = ["OVO", "CLIPGenerator", "MaskGenerator", "Instance3D", "RegionEncoder", "EncoderConfig", "SemanticMap"]


def __getattr__(name):          # lazy: importing the package must not need torch/CUDA (used by the ABI tests)
    if name == "OVO":
        from .ovo import OVO
        return OVO
    if name == "CLIPGenerator":
        from .clip_generator import CLIPGenerator
        return CLIPGenerator
    if name == "MaskGenerator":
        from .mask_generator import MaskGenerator
        return MaskGenerator
    if name in ("RegionEncoder", "EncoderConfig"):
        from . import encoder
        return getattr(encoder, name)
    if name == "SemanticMap":
        from .map import SemanticMap
        return SemanticMap
    raise AttributeError(name)
