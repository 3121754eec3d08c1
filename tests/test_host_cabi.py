"""CPU suite, part 2: the drop-in boundary without a GPU.

 * librtiow_cuda.so loads and exports exactly the symbols include/rtiow_cuda.h declares;
 * host-only entry points (Camera::new, params defaults, tile sizes, the seeded scene builder) against the oracle;
 * error behaviour: no device -> NO_DEVICE (there is no CPU fallback), bad arguments -> INVALID_ARG, never a crash;
 * the product package never imports the oracle.
"""
import ctypes as C
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

from conftest import ROOT, final_camera

HEADER = ROOT / "include" / "rtiow_cuda.h"


def header_symbols():
    txt = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(rtiow_[a-z0-9_]+)\s*\(", txt)))


def test_so_exports_every_declared_symbol(capi):
    declared = header_symbols()
    assert declared == sorted(capi.SYMBOLS), "capi.SYMBOLS and include/rtiow_cuda.h disagree"
    out = subprocess.run(["nm", "-D", "--defined-only", str(capi.lib_path())], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    missing = [s for s in declared if s not in exported]
    assert not missing, f"librtiow_cuda.so does not export {missing}"
    L = capi.lib()
    for s in declared:
        assert getattr(L, s) is not None
    assert L.rtiow_abi_version() == capi.ABI_VERSION


def test_so_contains_sm100a_code_only(capi):
    out = subprocess.run(["cuobjdump", "-lelf", str(capi.lib_path())], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, f"expected sm_100a only, found {archs}"


def test_struct_layouts_match_header(capi):
    # sizes the Rust -sys crate / C callers rely on (repr(C))
    assert C.sizeof(capi.Camera) == 22 * 8
    assert C.sizeof(capi.Params) == 48 and capi.Params.t_min.offset == 16 and capi.Params.seed.offset == 24 and capi.Params.tile_rows.offset == 40
    assert C.sizeof(capi.Stats) == 72 and capi.Stats.scan_backend.offset == 64
    assert C.sizeof(capi.Spheres) == 48 and C.sizeof(capi.Materials) == 48


def test_params_default_mirror_reference_constants(capi):
    p = capi.default_params()
    assert (p.width, p.height, p.spp, p.max_depth, p.alpha) == (200, 133, 100, 50, 255)       # main.rs:24-28,137
    assert p.t_min == 0.0001 and p.precision == capi.F32 and p.tile_rows >= 1                 # main.rs:44


def test_camera_new_matches_oracle_exactly(capi, oracle):
    for args in [((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 1.5, 0.1, 10.0), ((-2, 2, 1), (0, 0, -1), (0, 1, 0), 90.0, 16 / 9, 2.0, 3.4),
                 ((3, 3, 2), (0, 0, -1), (0.1, 1, 0.2), 37.5, 2.0, 0.0, 5.2)]:
        g, o = capi.camera_new(*args), oracle.camera_new(*args)
        for f in ("origin", "lower_left_corner", "horizontal", "vertical", "u", "v", "w"):
            assert list(getattr(g, f)) == list(getattr(o, f).np()), f
        assert g.lens_radius == o.lens_radius


def test_random_scene_follows_main_rs(capi):
    s = capi.random_scene(1)
    n = len(s["radius"])
    assert 500 <= n <= 533                                                                     # 1 + <=529 + 3 (main.rs:64-99)
    assert list(s["center"][0]) == [0, -1000, 0] and s["radius"][0] == 1000 and s["mat_kind"][0] == capi.MAT_LAMBERTIAN
    assert np.allclose(s["center"][-3:], [[0, 1, 0], [-4, 1, 0], [4, 1, 0]]) and list(s["radius"][-3:]) == [1, 1, 1]
    assert list(s["mat_kind"][-3:]) == [capi.MAT_DIELECTRIC, capi.MAT_LAMBERTIAN, capi.MAT_METAL] and s["mat_param"][-1] == 0.0
    small = slice(1, n - 3)
    c = s["center"][small]
    assert (s["radius"][small] == 0.2).all() and (c[:, 1] == 0.2).all()
    assert (np.linalg.norm(c - [4, 0.2, 0], axis=1) > 0.9).all()                               # main.rs:72
    a, b = np.floor(c[:, 0]), np.floor(c[:, 2])
    assert a.min() == -11 and a.max() == 11 and b.min() == -11 and b.max() == 11               # inclusive grid -11..=11
    assert ((c[:, 0] - a) < 0.9).all() and ((c[:, 2] - b) < 0.9).all()
    k = s["mat_kind"][small]
    frac = np.bincount(k, minlength=3) / len(k)
    assert abs(frac[0] - 0.80) < 0.06 and abs(frac[1] - 0.15) < 0.05 and abs(frac[2] - 0.05) < 0.04   # main.rs:74,78,83
    metal = s["mat_param"][small][k == capi.MAT_METAL]
    assert (0 <= metal).all() and (metal < 0.5).all() and (s["mat_albedo"][small][k == capi.MAT_METAL] >= 0.5).all()
    assert (s["mat_param"][small][k == capi.MAT_DIELECTRIC] == 1.5).all()
    # seeded: reproducible, and different seeds differ
    s2, s3 = capi.random_scene(1), capi.random_scene(2)
    assert all(np.array_equal(s[key], s2[key]) for key in s) and not np.array_equal(s["center"][1:20], s3["center"][1:20])


@pytest.mark.parametrize("mode,kind", [(1, 0), (2, 1), (3, 2)])
def test_material_isolation_scenes(capi, mode, kind):
    s = capi.random_scene(2, 11, mode)
    assert (s["mat_kind"][1:] == kind).all() and s["mat_kind"][0] == capi.MAT_LAMBERTIAN       # ground stays diffuse
    if mode == 3:
        assert (s["radius"] == -0.9).sum() == 1                                                # hollow shell (negative radius)


def test_big_scene_size(capi):
    s = capi.random_scene(3, 50, 0)
    assert 10_000 < len(s["radius"]) <= 101 * 101 + 4                                          # BASELINE configs[3]


def test_tile_buffer_bytes_and_partition(capi):
    from rtiow_b200 import partition as pr
    L = capi.lib()
    for (W, H, T, G) in [(1200, 675, 4, 8), (1200, 675, 4, 1), (3840, 2160, 8, 8), (400, 225, 1, 3), (201, 133, 7, 4), (8, 2, 64, 8)]:
        p = capi.default_params(width=W, height=H, tile_rows=T)
        n = C.c_size_t(0)
        assert L.rtiow_tile_buffer_bytes(C.byref(p), G, C.byref(n)) == capi.OK
        rows = [pr.rows_of_rank(H, T, G, r) for r in range(G)]
        assert sum(len(r) for r in rows) == H and sorted(sum(rows, [])) == list(range(H))     # a partition of the frame
        assert n.value == pr.max_rows_per_rank(H, T, G) * W * 4 and n.value >= max(len(r) for r in rows) * W * 4


def test_no_device_is_an_error_not_a_fallback(capi):
    if capi.device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(capi.RtiowError) as e:
        capi.Context(1)
    assert e.value.code == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)
    with pytest.raises(capi.RtiowError):
        capi.Context(device=0)


def test_invalid_arguments_do_not_crash(capi):
    L = capi.lib()
    assert L.rtiow_ctx_create(0, None) == capi.ERR_INVALID_ARG
    h = C.c_void_p()
    assert L.rtiow_ctx_create(0, C.byref(h)) == capi.ERR_INVALID_ARG and b"n_gpus" in L.rtiow_last_error()
    assert L.rtiow_scene_upload(None, None, None) == capi.ERR_INVALID_ARG
    assert L.rtiow_render(None, None, None, None, None) == capi.ERR_INVALID_ARG
    assert L.rtiow_render_progressive(None, None, None, 4, capi.PROGRESS_FN(0), None, None, None) == capi.ERR_INVALID_ARG
    assert L.rtiow_camera_new(None, None, None, 1.0, 1.0, 1.0, 1.0, None) == capi.ERR_INVALID_ARG
    assert L.rtiow_tile_buffer_bytes(None, 1, None) == capi.ERR_INVALID_ARG
    assert L.rtiow_render_to_frame_device(None, None, None, 0, 1, None, None, None) == capi.ERR_INVALID_ARG
    n = C.c_size_t(0)
    for bad in (dict(width=1), dict(height=0), dict(spp=0), dict(tile_rows=0), dict(precision=7), dict(t_min=-1.0)):
        p = capi.default_params(**bad)
        assert L.rtiow_tile_buffer_bytes(C.byref(p), 2, C.byref(n)) == capi.ERR_INVALID_ARG
    assert L.rtiow_random_scene(1, 11, 9, 10, None, None, None, None, None, None, None, None) == capi.ERR_INVALID_ARG
    L.rtiow_ctx_destroy(None)                                                                   # no-op, must not crash


def test_api_mirror_describes_scene(capi):
    import rtiow_b200 as r
    world = r.HittableList()
    ground = r.Lambertian((0.5, 0.5, 0.5))
    world.push(r.Sphere((0, -1000, 0), 1000, ground))                                          # main.rs:62-64
    world.push(r.Sphere((0, 1, 0), 1.0, r.Dialectric(1.5)))
    world.push(r.Sphere((4, 1, 0), 1.0, r.Metal((0.7, 0.6, 0.5), 0.0)))
    world.push(r.Sphere((5, 1, 0), -0.5, ground))
    a = world.to_arrays()
    assert list(a["mat_index"]) == [0, 1, 2, 0] and list(a["mat_kind"]) == [0, 2, 1] and list(a["radius"]) == [1000, 1, 1, -0.5]
    assert list(a["mat_param"]) == [0, 1.5, 0.0] and r.Dielectric is r.Dialectric
    world.push("a triangle")
    with pytest.raises(capi.RtiowError) as e:
        world.to_arrays()
    assert e.value.code == capi.ERR_UNSUPPORTED
    assert len(r.random_scene(1)) == len(capi.random_scene(1)["radius"])


def test_product_never_touches_the_oracle():
    """the product path must not import, link or execute anything under oracle/"""
    offenders = []
    for f in list((ROOT / "rtiow_b200").rglob("*.py")) + list((ROOT / "rtiow_b200").rglob("*.cu*")) + list((ROOT / "rtiow_b200").rglob("*.[ch]pp")) \
            + [ROOT / "include" / "rtiow_cuda.h"]:
        t = f.read_text()
        if re.search(r"(from|import)\s+oracle|rtiow_oracle|librtiow_oracle|oracle/", t):
            offenders.append(str(f.relative_to(ROOT)))
    assert not offenders, offenders
    out = subprocess.run(["ldd", str(ROOT / "rtiow_b200" / "lib" / "librtiow_cuda.so")], capture_output=True, text=True).stdout
    assert "oracle" not in out


def test_bench_reference_arm_and_configs():
    """bench.py: every BASELINE configuration is selectable, and the reference arm (the CPU restatement, no GPU) prints one JSON
    line of the bench contract with impl = reference"""
    import json, subprocess, sys
    root = Path(__file__).resolve().parent.parent
    sys.path.insert(0, str(root))
    import bench
    assert set(bench.CONFIGS) == {"cfg1", "cfg2", "cfg3-lambertian", "cfg3-metal", "cfg3-dielectric", "cfg4", "cfg5"}
    assert bench.CONFIGS["cfg2"][1:4] == (1200, 675, 500) and bench.CONFIGS["cfg5"][1:4] == (3840, 2160, 1024)
    r = subprocess.run([sys.executable, str(root / "bench.py"), "--impl", "reference", "--config", "cfg1", "--steps", "1", "--warmup", "0",
                        "--cpu-sample-spp", "1"], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr
    d = json.loads(r.stdout.strip().splitlines()[-1])
    assert d["impl"] == "reference" and d["metric"] == "Mpaths/s" and d["value"] > 0 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] == "port" and d["e2e"]["h2d_bytes_per_step"] == 0 and d["config"]["width"] == 400
