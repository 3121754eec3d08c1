"""Kernel-level throughput of every BASELINE.json configuration on one GPU (not the bench contract: see bench.py).
    python tools/bench_configs.py [reps]   -> one JSON line per config (written to stdout)"""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtiow_b200 import capi

reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
CFG = [("cfg1 final 400x225@10", 1, 11, 0, 400, 225, 10), ("cfg2 final 1200x675@500", 1, 11, 0, 1200, 675, 500),
       ("cfg3 all-Lambertian 800x450@100", 1, 11, 1, 800, 450, 100), ("cfg3 all-Metal 800x450@100", 1, 11, 2, 800, 450, 100),
       ("cfg3 all-Dialectric+shell 800x450@100", 1, 11, 3, 800, 450, 100), ("cfg4 10k spheres 1920x1080@256", 1, 50, 0, 1920, 1080, 256),
       ("cfg5 final 3840x2160@1024", 1, 11, 0, 3840, 2160, 1024)]
with capi.Context(1) as ctx:
    peak, _ = ctx.fp32_peak_probe(True, 200.0)
    for name, seed, g, mode, W, H, spp in CFG:
        sc = capi.random_scene(seed, g, mode)
        ctx.upload_scene(**sc)
        cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
        prm = capi.default_params(width=W, height=H, spp=spp, seed=1)
        best = None
        for _ in range(reps):
            _, st = ctx.render(cam, prm)
            if best is None or st["kernel_ms"] < best["kernel_ms"]:
                best = st
        tf = best["sphere_tests"] * 17 / best["kernel_ms"] / 1e9
        print(json.dumps({"config": name, "n_spheres": len(sc["radius"]), "paths": best["paths"], "kernel_ms": round(best["kernel_ms"], 3),
                          "mpaths_s": round(best["paths"] / best["kernel_ms"] / 1e3, 1), "rays_per_path": round(best["rays_traced"] / best["paths"], 4),
                          "tflops_17": round(tf, 2), "frac_of_fp32_peak": round(tf / peak, 4), "fp32_peak_tflops": round(peak, 2)}), flush=True)
