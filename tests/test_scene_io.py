"""CPU suite: seeded scenes and scene files — main.rs:59-102's world as data.

The same C++ source (rtiow_b200/csrc/scene_gen.cpp) is compiled into librtiow_cuda.so and, host-only, into
oracle/build/librtiow_scene.so (what bench.py's CPU arms and the reference arm load, so they never map the CUDA library);
both must produce the same arrays, and a scene must survive the text file bit for bit."""
from pathlib import Path

import numpy as np
import pytest

GOLDEN = Path(__file__).resolve().parent / "golden" / "scene_seed1_grid2.txt"


def _same(a, b):
    return all(np.array_equal(a[k], b[k]) for k in ("center", "radius", "mat_kind", "mat_albedo", "mat_param"))


@pytest.mark.parametrize("seed,grid,mode", [(1, 11, 0), (2, 11, 0), (1, 11, 3), (7, 3, 2), (1, 50, 0)])
def test_host_only_scene_library_equals_the_product_library(capi, oracle, seed, grid, mode):
    a, b = capi.random_scene(seed, grid, mode), oracle.random_scene(seed, grid, mode)
    assert _same(a, b) and len(a["radius"]) > 4


def test_scene_file_round_trip_is_bit_exact(capi, oracle, tmp_path):
    sc = capi.random_scene(3, 11, 0)
    sc["radius"][5] = -0.125; sc["center"][6] = [1e-300, -1.0 / 3.0, 12345.678901234567]       # negative radius, awkward doubles
    f = tmp_path / "world.txt"
    capi.save_scene(f, **sc)
    assert _same(sc, capi.load_scene(f)) and _same(sc, oracle.load_scene(f))
    assert f.read_text().splitlines()[:2] == ["rtiow-scene 1", f"n {len(sc['radius'])}"]


def test_scene_file_shares_materials_through_mat_index(capi, tmp_path):
    f = tmp_path / "shared.txt"
    capi.save_scene(f, center=[[0, 0, 0], [1, 0, 0], [2, 0, 0]], radius=[1, 1, 1], mat_index=[1, 0, 1], mat_kind=[0, 2],
                    mat_albedo=[[.1, .2, .3], [1, 1, 1]], mat_param=[0, 1.5])
    sc = capi.load_scene(f)
    assert sc["mat_kind"].tolist() == [2, 0, 2] and sc["mat_param"].tolist() == [1.5, 0, 1.5] and sc["mat_index"].tolist() == [0, 1, 2]


def test_scene_file_errors(capi, tmp_path):
    import ctypes as C
    L = capi.lib()
    n = C.c_uint32(0)
    assert L.rtiow_scene_load(str(tmp_path / "missing.txt").encode(), 0, None, None, None, None, None, None, None, C.byref(n)) == capi.ERR_INVALID_ARG
    bad = tmp_path / "bad.txt"; bad.write_text("rtiow-scene 9\nn 1\n0 0 0 1 0 1 1 1 0\n")
    assert L.rtiow_scene_load(str(bad).encode(), 0, None, None, None, None, None, None, None, C.byref(n)) == capi.ERR_UNSUPPORTED
    trunc = tmp_path / "trunc.txt"; trunc.write_text("rtiow-scene 1\nn 2\n0 0 0 1 0 1 1 1 0\n")
    with pytest.raises(capi.RtiowError):
        capi.load_scene(trunc)
    kind = tmp_path / "kind.txt"; kind.write_text("rtiow-scene 1\nn 1\n0 0 0 1 5 1 1 1 0\n")
    with pytest.raises(capi.RtiowError):
        capi.load_scene(kind)
    ok = tmp_path / "ok.txt"; capi.save_scene(ok, **capi.random_scene(1, 2, 0))
    a = np.zeros(3); k = np.zeros(3, np.uint32); alb = np.zeros(9)
    p = lambda x: x.ctypes.data_as(C.c_void_p)
    assert L.rtiow_scene_load(str(ok).encode(), 3, p(a), p(a), p(a), p(a), p(k), p(alb), p(a), C.byref(n)) == capi.ERR_NOMEM and n.value > 3
    with pytest.raises(capi.RtiowError):
        capi.save_scene(tmp_path / "nodir" / "x.txt", **capi.random_scene(1, 2, 0))


def test_golden_scene_file(capi, tmp_path):
    """the committed file pins the generator's stream AND the file format: tests/golden/scene_seed1_grid2.txt = random_scene(1, 2, 0)"""
    f = tmp_path / "g.txt"
    capi.save_scene(f, **capi.random_scene(1, 2, 0))
    assert f.read_text() == GOLDEN.read_text()
    sc = capi.load_scene(GOLDEN)
    assert sc["center"][0].tolist() == [0.0, -1000.0, 0.0] and sc["radius"][0] == 1000.0          # main.rs:63-64
    assert sc["center"][-1].tolist() == [4.0, 1.0, 0.0] and sc["mat_kind"][-1] == capi.MAT_METAL   # main.rs:98-99
