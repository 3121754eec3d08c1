"""Diagnostic (GPU): where do the FP32 and the tensor-core filter disagree?  Not part of the product path.

  python tools/diag_tensor.py [n_rays]

Shoots render-like rays (camera rays, bounce rays leaving the ground and the sphere surfaces) through rtiow_hitlist_batch on
both backends and prints every ray whose hit differs, with the f64 discriminant of the sphere one side found and the other
missed.  Then renders a small frame on both backends and lists the pixels that differ.
"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtiow_b200 import capi  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000
    rng = np.random.default_rng(3)
    scene = capi.random_scene(1, 11, 0)
    C, R = scene["center"], scene["radius"]
    ctx = capi.Context(1)
    ctx.upload_scene(**scene)
    k = n // 4
    o = np.empty((n, 3)); d = np.empty((n, 3))
    o[:k] = (13, 2, 3) + 0.05 * rng.standard_normal((k, 3))
    d[:k] = np.stack([rng.uniform(-11, 11, k), rng.uniform(-0.5, 2.5, k), rng.uniform(-11, 11, k)], 1) - o[:k]
    # rays leaving sphere surfaces (small spheres and the ground)
    sid = rng.integers(0, len(R), n - k)
    nrm = rng.standard_normal((n - k, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    grd = sid == 0
    nrm[grd] = np.stack([rng.uniform(-0.03, 0.03, grd.sum()), np.ones(grd.sum()), rng.uniform(-0.03, 0.03, grd.sum())], 1)
    nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
    o[k:] = C[sid] + nrm * np.abs(R[sid])[:, None]
    dd = rng.standard_normal((n - k, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
    d[k:] = nrm + dd
    o, d = o.astype(np.float32).astype(float), d.astype(np.float32).astype(float)
    ctx.set_scan_backend(capi.SCAN_FP32); a = ctx.hitlist_batch(o, d)
    ctx.set_scan_backend(capi.SCAN_TENSOR); b = ctx.hitlist_batch(o, d)
    bad = np.nonzero((a["index"] != b["index"]) | (a["hit"] != b["hit"]))[0]
    print(f"hitlist: {n} rays, hit rate {a['hit'].mean():.3f}, {len(bad)} differ")
    for i in bad[:40]:
        ia, ib = int(a["index"][i]), int(b["index"][i])
        line = f"  ray {i}: fp32 idx {ia} t {a['t'][i]:.6g} | tensor idx {ib} t {b['t'][i]:.6g} | o {o[i]} d {d[i]}"
        for j in {ia, ib} - {-1}:
            dh = d[i] / np.linalg.norm(d[i]); oc = C[j] - o[i]; hb = oc @ dh
            line += f" | sphere {j} r {R[j]} disc {hb * hb - (oc @ oc - R[j] ** 2):.3e} tca {hb:.4g}"
        print(line)
    for W, H, spp in ((400, 225, 10), (64, 36, 33)):
        cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
        prm = capi.default_params(width=W, height=H, spp=spp, seed=7)
        ctx.set_scan_backend(capi.SCAN_FP32); ia, sa = ctx.render(cam, prm)
        imgs = []
        for rep in range(3):
            ctx.set_scan_backend(capi.SCAN_TENSOR); ib, sb = ctx.render(cam, prm)
            imgs.append((ib.copy(), sb["rays_traced"]))
        print(f"render {W}x{H}@{spp}: fp32 rays {sa['rays_traced']}, tensor rays {[r for _, r in imgs]}, "
              f"pixels differing from fp32 {[int((im != ia).any(axis=2).sum()) for im, _ in imgs]}, "
              f"tensor runs identical to each other: {all(np.array_equal(imgs[0][0], im) for im, _ in imgs)}")
        ys, xs = np.nonzero((imgs[0][0] != ia).any(axis=2))
        print("   first differing pixels (y, x):", list(zip(ys.tolist(), xs.tolist()))[:12])


if __name__ == "__main__" and not (len(sys.argv) > 2 and sys.argv[2] in ("bounce", "pixel")):
    main()


def bounce_diag(n=1_000_000, W=400, H=225):
    """real bounce rays: camera rays advanced with the oracle's scatter; both backends on the same f32-representable rays"""
    from oracle import oracle as o
    scene = capi.random_scene(1, 11, 0); sc = o.Scene(**scene)
    C, R = scene["center"], scene["radius"]
    ctx = capi.Context(1); ctx.upload_scene(**scene)
    ocam = o.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    rng = np.random.default_rng(0)
    r = o.get_ray_batch(ocam, rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(-0.7, 0.7, (n, 2)))
    orig, d = r["orig"], r["dir"]
    for bounce in range(10):
        o32 = orig.astype(np.float32).astype(np.float64); d32 = d.astype(np.float32).astype(np.float64)
        ctx.set_scan_backend(capi.SCAN_FP32); a = ctx.hitlist_batch(o32, d32, 1e-4)
        ctx.set_scan_backend(capi.SCAN_TENSOR); b = ctx.hitlist_batch(o32, d32, 1e-4)
        bad = np.nonzero(a["index"] != b["index"])[0]
        print(f"bounce {bounce}: rays {len(o32)}, hit frac {a['hit'].mean():.3f}, max |o| {np.abs(o32).max():.0f}, backends differ on {len(bad)}")
        for i in bad[:10]:
            ia, ib = int(a["index"][i]), int(b["index"][i])
            line = f"   o={o32[i]} d={d32[i]} fp32 idx {ia} t {a['t'][i]:.6g} | tensor idx {ib} t {b['t'][i]:.6g}"
            for j in {ia, ib} - {-1}:
                dh = d32[i] / np.linalg.norm(d32[i]); oc = C[j] - o32[i]; hb = oc @ dh
                line += f" | sphere {j} c {C[j]} r {R[j]} disc {hb * hb - (oc @ oc - R[j] ** 2):.3e} tca {hb:.5g}"
            print(line)
        ref = o.world_hit_batch(sc, o32, d32)
        hm = ref["hit"] == 1
        k = ref["index"][hm]
        smp = rng.normal(size=(hm.sum(), 3)); smp *= (rng.uniform(0, 1, (hm.sum(), 1)) ** (1 / 3)) / np.linalg.norm(smp, axis=1, keepdims=True)
        kinds = scene["mat_kind"][k]
        smp[kinds == 2, 0] = rng.uniform(0, 1, (kinds == 2).sum())
        s = o.scatter_batch(kinds, scene["mat_albedo"][k], scene["mat_param"][k], o32[hm], d32[hm], ref["p"][hm], ref["normal"][hm], ref["front_face"][hm], smp)
        keep = s["some"] == 1
        orig, d = s["orig"][keep], s["dir"][keep]
        if len(orig) < 1000:
            break


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "bounce":
    bounce_diag()


def pixel_diag(y=32, x=188, W=400, H=225, spp=10, seed=7):
    """re-trace the samples of one pixel through rtiow_ray_color_trace_batch on both backends and print where the paths part"""
    scene = capi.random_scene(1, 11, 0)
    C, R = scene["center"], scene["radius"]
    ctx = capi.Context(1); ctx.upload_scene(**scene)
    cam = capi.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
    pj = H - 1 - y
    pix = np.full(spp, pj * W + x, np.uint32); smp = np.arange(spp, dtype=np.uint32)
    u = ctx.sampler_batch(pix, smp, np.zeros(spp, np.uint32), seed)
    f = np.float32
    su = (f(x) + u[:, 0].astype(f)) * f(1.0 / (W - 1)); sv = (f(pj) + u[:, 1].astype(f)) * f(1.0 / (H - 1))
    r = ctx.get_ray_batch(cam, su.astype(float), sv.astype(float), u[:, 4:6])
    out = {}
    for name, be in (("fp32", capi.SCAN_FP32), ("tensor", capi.SCAN_TENSOR)):
        ctx.set_scan_backend(be)
        out[name] = ctx.ray_color_trace_batch(r["orig"], r["dir"], pix, smp, seed)
    a, b = out["fp32"], out["tensor"]
    print("rays per sample fp32  ", a["rays"].tolist())
    print("rays per sample tensor", b["rays"].tolist())
    for s in range(spp):
        ia, ib = a["index"][s], b["index"][s]
        if np.array_equal(ia, ib):
            continue
        k = int(np.nonzero(ia != ib)[0][0])
        kr = int(np.nonzero((a["ray"][s] != b["ray"][s]).any(axis=1))[0][0])
        print(f"sample {s}: first ray whose (o, d) differ: {kr} (hit before it: {ia[kr - 1] if kr else None}); |delta o| {np.abs(a['ray'][s, kr, :3] - b['ray'][s, kr, :3]).max():.3e} "
              f"|delta d| {np.abs(a['ray'][s, kr, 3:] - b['ray'][s, kr, 3:]).max():.3e}")
        print(f"   fp32   ray {kr - 1}: {a['ray'][s, kr - 1].tolist()}\n   tensor ray {kr - 1}: {b['ray'][s, kr - 1].tolist()}")
        print(f"   fp32   ray {kr}: {a['ray'][s, kr].tolist()}\n   tensor ray {kr}: {b['ray'][s, kr].tolist()}")
        ro, rd = a["ray"][s, k, :3], a["ray"][s, k, 3:]
        print(f"sample {s}: paths part at ray {k}: fp32 hits {ia[k]}, tensor hits {ib[k]}; previous hits {ia[:k].tolist()}")
        print(f"   ray o={ro.tolist()} d={rd.tolist()} (tensor's ray equal: {np.array_equal(a['ray'][s, k], b['ray'][s, k])})")
        for j in {int(ia[k]), int(ib[k])} - {-1}:
            oc = C[j] - ro; hb = oc @ rd; aa = rd @ rd
            print(f"   sphere {j}: c={C[j].tolist()} r={R[j]} disc {hb * hb - aa * (oc @ oc - R[j] ** 2):.4e} tca {hb:.6g} |oc| {np.linalg.norm(oc):.6g}")
        ctx.set_scan_backend(capi.SCAN_FP32); ha = ctx.hitlist_batch(ro[None], rd[None])
        ctx.set_scan_backend(capi.SCAN_TENSOR); hb_ = ctx.hitlist_batch(ro[None], rd[None])
        print(f"   the same ray through hitlist (no self sphere): fp32 {ha['index'][0]} t {ha['t'][0]:.6g} | tensor {hb_['index'][0]} t {hb_['t'][0]:.6g}")


if __name__ == "__main__" and len(sys.argv) > 2 and sys.argv[2] == "pixel":
    pixel_diag()
