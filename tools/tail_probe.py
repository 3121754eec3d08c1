"""Fixed cost of a frame: fit kernel_ms = a + b * spp at several max_depth (is the tail made of the longest paths?)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtiow_b200 import capi
W, H = 1200, 675
with capi.Context(1) as ctx:
    ctx.upload_scene(**capi.random_scene(1))
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    for depth in (50, 10, 3, 1):
        t = {}
        for spp in (20, 40, 80, 160):
            prm = capi.default_params(width=W, height=H, spp=spp, seed=1, max_depth=depth)
            best = 1e9
            for _ in range(3):
                _, st = ctx.render(cam, prm)
                best = min(best, st["kernel_ms"])
            t[spp] = best
        b = (t[160] - t[40]) / 120; a = t[40] - 40 * b
        print(f"depth {depth}: " + " ".join(f"{s}spp {v:.3f}ms" for s, v in t.items()) + f" | fit a = {a:.3f} ms, b = {b:.4f} ms/spp", flush=True)
