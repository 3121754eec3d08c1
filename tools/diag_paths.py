"""GPU diagnostic: where do f32 closest-hit results differ from the f64 oracle on real bounce rays?"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from rtiow_b200 import capi
from oracle import oracle as o

scene = capi.random_scene(1); sc = o.Scene(**scene)
ctx = capi.Context(1); ctx.upload_scene(**scene)
W, H = 400, 225
ocam = o.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, W / H, 0.1, 10.0)
rng = np.random.default_rng(0)
n = 200000
r = o.get_ray_batch(ocam, rng.uniform(0, 1, n), rng.uniform(0, 1, n), rng.uniform(-0.7, 0.7, (n, 2)))
orig, d = r["orig"], r["dir"]
L = o.lib()
for bounce in range(6):
    o32 = orig.astype(np.float32).astype(np.float64); d32 = d.astype(np.float32).astype(np.float64)
    ref = o.world_hit_batch(sc, o32, d32)
    got = ctx.hitlist_batch(o32, d32, 1e-4)
    want = np.where(ref["hit"] == 1, ref["index"], -1)
    bad = got["index"] != want
    print(f"bounce {bounce}: rays {len(o32)}, hit frac {np.mean(want>=0):.3f}, disagree {bad.sum()} ({bad.mean():.5%})")
    for i in np.nonzero(bad)[0][:12]:
        kind_w = scene["mat_kind"][want[i]] if want[i] >= 0 else -1
        print(f"   ray o={o32[i]} |d|={np.linalg.norm(d32[i]):.4f} d^={d32[i]/np.linalg.norm(d32[i])} want idx {want[i]} t {ref['t'][i]:.6g} kind {kind_w} | got idx {got['index'][i]} t {got['t'][i]:.6g}")
    # classify: got hit where oracle missed / different sphere / got miss
    print("   got-hit-but-oracle-miss", int(((got['index']>=0)&(want<0)).sum()), " got-miss-but-oracle-hit", int(((got['index']<0)&(want>=0)).sum()),
          " different-sphere", int(((got['index']>=0)&(want>=0)&bad).sum()))
    m = ~bad & (want >= 0)
    te = np.abs(got["t"][m] - ref["t"][m]) / np.abs(ref["t"][m])
    print(f"   t rel err: max {te.max():.3g}  p99.9 {np.percentile(te,99.9):.3g}")
    # advance surviving rays with the oracle scatter (fresh random samples)
    k = want[want >= 0]
    hm = want >= 0
    smp = rng.normal(size=(hm.sum(), 3)); smp *= (rng.uniform(0, 1, (hm.sum(), 1)) ** (1 / 3)) / np.linalg.norm(smp, axis=1, keepdims=True)
    kinds = scene["mat_kind"][k]
    smp[kinds == 2, 0] = rng.uniform(0, 1, (kinds == 2).sum())
    s = o.scatter_batch(kinds, scene["mat_albedo"][k], scene["mat_param"][k], o32[hm], d32[hm], ref["p"][hm], ref["normal"][hm], ref["front_face"][hm], smp)
    keep = s["some"] == 1
    orig, d = s["orig"][keep], s["dir"][keep]
