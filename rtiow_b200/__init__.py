"""rtiow_b200 — B200 (sm_100a) render backend for the hot path of Druthyn/rtiow.

The product is librtiow_cuda.so (rtiow_b200/csrc, C ABI in include/rtiow_cuda.h).  `capi` binds it with
ctypes; `api` mirrors the reference's caller-facing names.  Nothing here imports the CPU oracle.
"""
from . import capi  # noqa: F401
from .api import (Camera, Dialectric, Dielectric, HittableList, Lambertian, Metal, RenderParams, Sphere,  # noqa: F401
                  random_scene, render)
