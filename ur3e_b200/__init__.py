"""ur3e_b200: B200-native batched simulator for the UR3e + Robotiq 2F85 (+ mug) Gymnasium environments.

The hot path (controller -> physics substeps -> obs/reward/done) runs as one fused sm_100a kernel, one
environment per warp, behind a C ABI (include/ur3e_b200.h).  No CPU fallback exists.
"""
from . import _lib  # noqa: F401
from .model import Model, asset  # noqa: F401

__all__ = ["Model", "asset", "_lib"]
