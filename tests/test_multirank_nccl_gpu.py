"""GPU suite, >= 2 GPUs: the gather INSIDE librtiow_cuda.so.

One process per GPU (rtiow_ctx_create_rank + rtiow_render_rank: ncclAllGather, and the fused CUDA-IPC peer-store epilogue),
spawned by tools/multirank_check.py with a file as the only rendezvous — no torch.distributed on the data path — and one
process driving n GPUs (rtiow_ctx_create(n)) with each of its gathers.  Every frame must equal the single-GPU frame byte for
byte (pixels are keyed (pixel, sample, bounce), accumulation is integer: main.rs:122-139's image does not depend on who
rendered which row).  Skipped on a 1-GPU box; `gpurun --gpus 2` runs them (log kept under profiles/).
"""
import json
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

from conftest import final_camera

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _need(capi, n):
    if capi.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("world", [2, 4, 8])
def test_rank_processes_gather_inside_the_library(capi, world):
    _need(capi, world)
    r = subprocess.run([sys.executable, str(ROOT / "tools" / "multirank_check.py"), "--world", str(world), "--width", "403", "--height", "227", "--spp", "6"],
                       capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith("{")][-1])
    assert res["nccl_identical_to_1gpu"] and res["auto_identical_to_1gpu"] and res["rank1_frame_identical"], res
    assert "ncclAllGather" in res["nccl_note"] and res["enqueue_nccl_identical_to_1gpu"]
    if res.get("fused") != "unsupported":
        assert res["fused_identical_to_1gpu"] and res["enqueue_fused_identical_to_1gpu"] and "rank 0's frame" in res["fused_note"], res


@pytest.mark.parametrize("n", [2, 4, 8])
def test_one_process_n_gpus_every_gather(capi, final_scene, n):
    _need(capi, n)
    W, H, spp = 401, 226, 5
    cam = final_camera(capi, W / H)
    prm = capi.default_params(width=W, height=H, spp=spp, seed=3, tile_rows=2)
    with capi.Context(1) as one:
        one.upload_scene(**final_scene[0])
        ref, st1 = one.render(cam, prm)
    with capi.Context(n) as ctx:
        ctx.upload_scene(**final_scene[0])
        for mode, word in ((capi.GATHER_AUTO, "GPUs"), (capi.GATHER_NCCL, "ncclAllGather"), (capi.GATHER_FUSED, "peer memory")):
            ctx.set_gather(mode)
            img, st = ctx.render(cam, prm)
            assert np.array_equal(img, ref), f"gather mode {mode}: {(img != ref).sum()} bytes differ from the single-GPU frame"
            assert st["n_gpus"] == n and st["rays_traced"] == st1["rays_traced"] and word in ctx.gather_info()


def test_rank_ctx_world_1_needs_no_nccl(capi, final_scene):
    _need(capi, 1)
    cam = final_camera(capi, 16 / 9)
    prm = capi.default_params(width=160, height=90, spp=4, seed=2)
    with capi.Context(device=0, rank=0, world=1) as rc, capi.Context(1) as one:
        rc.upload_scene(**final_scene[0]); one.upload_scene(**final_scene[0])
        a, _ = rc.render_rank(cam, prm); b, _ = one.render(cam, prm)
        assert np.array_equal(a, b)
        ptr, st = rc.render_rank_device(cam, prm)
        assert ptr != 0 and st["n_gpus"] == 1


def test_enqueued_frames_equal_the_synchronous_call(capi, final_scene):
    """rtiow_render_rank_enqueue x K + rtiow_ctx_synchronize: frames back to back without a host round trip, same bytes, per-frame kernel times"""
    _need(capi, 1)
    W, H = 200, 113
    cam = final_camera(capi, W / H)
    with capi.Context(device=0, rank=0, world=1) as rc:
        rc.upload_scene(**final_scene[0])
        prms = [capi.default_params(width=W, height=H, spp=3, seed=s) for s in (5, 6, 7)]
        refs = [rc.render_rank(cam, p)[0] for p in prms]
        ptrs = [rc.render_rank_enqueue(cam, p) for p in prms]
        st = rc.synchronize()
        assert all(ptrs) and st["kernel_launches"] == 3 * 2 and st["kernel_ms"] > 0 and st["n_gpus"] == 1
        last = capi.device_to_host(ptrs[-1], W * H * 4).reshape(H, W, 4)
        assert np.array_equal(last, refs[-1])
        assert st["rays_traced"] == rc.render_rank(cam, prms[-1])[1]["rays_traced"]
        assert rc.synchronize()["kernel_launches"] == 0                      # nothing enqueued since
        again, _ = rc.render_rank(cam, prms[0])
        assert np.array_equal(again, refs[0])


def test_page_locked_caller_buffer_takes_the_frame_directly(capi, final_scene):
    """a pinned out_rgba (cudaHostAlloc / torch pin_memory) receives the D2H copy without the staging memcpy: same bytes, both entry points"""
    _need(capi, 1)
    torch = pytest.importorskip("torch")
    W, H = 200, 113
    cam = final_camera(capi, W / H)
    prm = capi.default_params(width=W, height=H, spp=3, seed=9)
    pinned = torch.zeros((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()
    with capi.Context(device=0, rank=0, world=1) as rc, capi.Context(1) as one:
        rc.upload_scene(**final_scene[0]); one.upload_scene(**final_scene[0])
        ref, _ = one.render(cam, prm)
        a, _ = rc.render_rank(cam, prm, out=pinned)
        assert a is pinned and np.array_equal(pinned, ref)
        pinned[:] = 0
        b, _ = one.render(cam, prm, out=pinned)
        assert np.array_equal(b, ref)

