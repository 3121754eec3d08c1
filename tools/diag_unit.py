"""print the worst t errors of the unit sphere-hit / hitlist tests"""
import sys
from pathlib import Path
import numpy as np
sys.path.insert(0, str(Path(__file__).resolve().parent.parent)); sys.path.insert(0, str(Path(__file__).resolve().parent.parent / "tests"))
from rtiow_b200 import capi
from oracle import oracle as o
from test_parity_unit_gpu import f32, t_err, not_grazing
ctx = capi.Context(1)
rng = np.random.default_rng(10)
n = 200_000
c = f32(rng.uniform(-8, 8, (n, 3))); r = f32(rng.uniform(0.1, 2.0, n) * rng.choice([1, 1, 1, -1], n))
oo = f32(rng.uniform(-14, 14, (n, 3)))
d = f32((c - oo) * rng.uniform(0.05, 2.0, (n, 1)) + rng.normal(size=(n, 3)) * np.abs(r)[:, None] * 0.8)
ref = o.sphere_hit_batch(c, r, oo, d, 1e-4, np.inf); got = ctx.sphere_hit_batch(c, r, oo, d, 1e-4, np.inf)
m = not_grazing(c, r, oo, d) & (ref["hit"] == 1) & (got["hit"] == 1)
te = t_err(got["t"], ref["t"], oo, d, c); te[~m] = 0
for i in np.argsort(-te)[:6]:
    oc = oo[i] - c[i]; a = d[i] @ d[i]; hb = oc @ d[i]; cc = oc @ oc - r[i] ** 2
    print(f"sphere_hit: te {te[i]:.3g} t_ref {ref['t'][i]:.8g} t_got {got['t'][i]:.8g} |oc| {np.linalg.norm(oc):.4g} r {r[i]:.4g} |d| {np.sqrt(a):.4g} disc/a/r2 {(hb*hb-a*cc)/a/r[i]**2:.4g} tca {-hb/np.sqrt(a):.5g}")
scene = capi.random_scene(1); scene = {k: (f32(v) if v.dtype == np.float64 else v) for k, v in scene.items()}; sc = o.Scene(**scene); ctx.upload_scene(**scene)
rng = np.random.default_rng(11)
ocam = o.camera_new((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0)
n1 = 60000
prim = o.get_ray_batch(ocam, rng.uniform(0, 1, n1), rng.uniform(0, 1, n1), rng.uniform(-0.7, 0.7, (n1, 2)))
n2 = 60000
k = rng.integers(1, sc.n, n2)
nrm = rng.normal(size=(n2, 3)); nrm /= np.linalg.norm(nrm, axis=1, keepdims=True)
so = scene["center"][k] + nrm * scene["radius"][k][:, None] * 1.05
sd = rng.normal(size=(n2, 3)) * rng.uniform(0.05, 2.0, (n2, 1))
go = np.stack([rng.uniform(-11, 11, n2 // 2), np.full(n2 // 2, 2e-3), rng.uniform(-11, 11, n2 // 2)], 1)
gd = rng.normal(size=(n2 // 2, 3)) * [1, 0.3, 1]
O = f32(np.concatenate([prim["orig"], so, go])); D = f32(np.concatenate([prim["dir"], sd, gd]))
ref = o.world_hit_batch(sc, O, D); got = ctx.hitlist_batch(O, D, 1e-4)
want = np.where(ref["hit"] == 1, ref["index"], -1)
m = (got["index"] == want) & (want >= 0)
hi = np.maximum(want, 0)
m &= not_grazing(scene["center"][hi], scene["radius"][hi], O, D) & (np.abs(ref["t"] - 1e-4) > 1e-5)
te = t_err(got["t"], ref["t"], O, D, scene["center"][hi]); te[~m] = 0
for i in np.argsort(-te)[:8]:
    print(f"hitlist: te {te[i]:.3g} ray#{i} idx {want[i]} r {scene['radius'][want[i]]} t_ref {ref['t'][i]:.8g} t_got {got['t'][i]:.8g} o {O[i]} |d| {np.linalg.norm(D[i]):.4g} dhat {D[i]/np.linalg.norm(D[i])}")
