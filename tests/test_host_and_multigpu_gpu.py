"""GPU suite: the compiled host mirror (C++ stand-in for the Rust caller) and the multi-GPU paths."""
import subprocess

import numpy as np
import pytest

from conftest import ROOT, final_camera

pytestmark = pytest.mark.gpu


def test_cpp_host_mirror_selftest():
    """rtiow_b200/host/rtiow.hpp: reference-named API -> describe() -> C ABI -> GPU; errors; PNG/PPM writers"""
    from rtiow_b200 import build
    exe = build.build_host_example()
    r = subprocess.run([str(exe), "--selftest"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "selftest ok" in r.stdout, r.stdout + r.stderr


def test_cpp_host_example_writes_png(tmp_path):
    from rtiow_b200 import build
    exe = build.build_host_example()
    out = tmp_path / "image.png"
    r = subprocess.run([str(exe), "--width", "120", "--spp", "4", "--out", str(out)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "Done." in r.stdout, r.stdout + r.stderr
    from PIL import Image
    im = np.array(Image.open(out))
    assert im.shape == (80, 120, 4) and (im[..., 3] == 255).all() and im[0, 0, 2] == 255     # 120/1.5 = 80 rows; sky on top


def test_single_process_multi_gpu_matches_single_gpu(capi, final_scene):
    """rtiow_ctx_create(n): interleaved row tiles on n GPUs gathered over NVLink peer copies == the 1-GPU bytes"""
    n = capi.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    cam = final_camera(capi, 16 / 9)
    prm = capi.default_params(width=400, height=225, spp=10, seed=1)
    with capi.Context(1) as c1:
        c1.upload_scene(**final_scene[0])
        base, st1 = c1.render(cam, prm)
    for g in sorted({2, n}):
        with capi.Context(g) as cg:
            cg.upload_scene(**final_scene[0])
            img, st = cg.render(cam, prm)
            seen = []
            imgp, stp = cg.render_progressive(cam, prm, 3, lambda k, n_, done, frame: seen.append(done) or False)
        assert np.array_equal(img, base), f"{g}-GPU image differs from the 1-GPU image"
        assert st["rays_traced"] == st1["rays_traced"] and st["n_gpus"] == g
        assert np.array_equal(imgp, base) and seen == [3, 6, 10] and stp["rays_traced"] == st1["rays_traced"], f"{g}-GPU progressive render differs"
