"""Where the end-to-end call spends its time beyond the kernel: rtiow_scene_upload, rtiow_render_rank into a page-locked and into a
pageable frame buffer, rtiow_render_rank_device (GPU box).    python tools/e2e_probe.py"""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from rtiow_b200 import capi
scene = capi.random_scene(1)
W, H = 1200, 675
with capi.Context(device=0, rank=0, world=1) as ctx:
    ctx.upload_scene(**scene)
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    prm = capi.default_params(width=W, height=H, spp=20, seed=1)
    out = torch.empty((H, W, 4), dtype=torch.uint8, pin_memory=True).numpy()
    out2 = np.empty((H, W, 4), np.uint8)
    for _ in range(3): ctx.upload_scene(**scene); ctx.render_rank(cam, prm, out=out)
    t0 = time.perf_counter()
    for _ in range(50): ctx.upload_scene(**scene)
    t1 = time.perf_counter()
    print(f"upload_scene: {(t1 - t0) / 50 * 1e3:.3f} ms")
    for name, o in (("pinned out", out), ("pageable out", out2)):
        ks = []; t0 = time.perf_counter()
        for _ in range(20):
            _, st = ctx.render_rank(cam, prm, out=o); ks.append(st["kernel_ms"])
        t1 = time.perf_counter()
        print(f"render_rank ({name}): {(t1 - t0) / 20 * 1e3:.3f} ms per call, kernel {np.mean(ks):.3f} ms -> overhead {(t1 - t0) / 20 * 1e3 - np.mean(ks):.3f} ms")
    ks = []; t0 = time.perf_counter()
    for _ in range(20):
        _, st = ctx.render_rank_device(cam, prm); ks.append(st["kernel_ms"])
    t1 = time.perf_counter()
    print(f"render_rank_device: {(t1 - t0) / 20 * 1e3:.3f} ms per call, kernel {np.mean(ks):.3f} ms -> overhead {(t1 - t0) / 20 * 1e3 - np.mean(ks):.3f} ms")
