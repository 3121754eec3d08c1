import sys
from pathlib import Path

import numpy as np
import pytest

# torch first, when it is there: it maps its own bundled libnccl.so.2, which librtiow_cuda.so then shares (as under bench.py).  The
# other order breaks `import torch` later in the same process: on a multi-GPU box the library's gathers dlopen the SYSTEM
# libnccl.so.2 (an older build), and the loader hands that one to libtorch_cuda.so, which needs symbols it does not have.
try:
    import torch  # noqa: F401
except Exception:  # the CPU-only suite and the library do not need it
    torch = None

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def oracle():
    """The CPU oracle (oracle/): test infrastructure, never on the product path."""
    from oracle import oracle as o
    o.lib()
    return o


@pytest.fixture(scope="session")
def capi():
    from rtiow_b200 import capi as c
    c.lib()
    return c


def _scene_pair(capi, oracle, seed=1, half_extent=11, mode=0):
    arrays = capi.random_scene(seed, half_extent, mode)
    return arrays, oracle.Scene(**arrays)


@pytest.fixture(scope="session")
def final_scene(capi, oracle):
    """RTIOW Part 1 final scene (main.rs:59-102), scene seed 1: (arrays for the C ABI, oracle.Scene)."""
    return _scene_pair(capi, oracle)


@pytest.fixture(scope="session")
def scene_factory(capi, oracle):
    return lambda seed=1, half_extent=11, mode=0: _scene_pair(capi, oracle, seed, half_extent, mode)


@pytest.fixture(scope="session")
def ctx(capi):
    """A single-GPU rtiow_ctx; only gpu-marked tests may request it."""
    if capi.device_count() == 0:
        pytest.skip("no CUDA device")
    c = capi.Context(1)
    yield c
    c.close()


@pytest.fixture()
def ctx_final(ctx, final_scene):
    """ctx with the final scene uploaded (re-uploaded per test: other tests replace the scene)."""
    ctx.upload_scene(**final_scene[0])
    return ctx


def final_camera(mod, aspect):
    """main.rs:108-118"""
    return mod.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, aspect, 0.1, 10.0)


def rel_err(a, b, floor=1e-6):
    a, b = np.asarray(a, float), np.asarray(b, float)
    return np.abs(a - b) / np.maximum(np.abs(b), floor)
