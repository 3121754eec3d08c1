"""CPU suite, part 1: pin the oracle.

The reference has no tests, so the oracle is pinned by
  (1) the one golden artefact the reference ships — sky-only pixels of rtiow_part1_final.png;
  (2) analytic known answers for every function on the path;
  (3) Random123 known-answer vectors for Philox4x32-10;
  (4) an independently written numpy restatement (tests/np_restatement.py);
  (5) distribution tests that the inversion samplers equal the reference's rejection samplers.
"""
import json
from pathlib import Path

import numpy as np
import pytest

import np_restatement as npr
from conftest import final_camera, rel_err

GOLDEN = Path(__file__).parent / "golden"


# --------------------------------------------------------------------------------------------- (1) PNG fixture
@pytest.mark.parametrize("sampler", ["direct", "rejection"])
def test_png_sky_rows(oracle, final_scene, sampler):
    """Camera::new + miss branch of ray_color + to_rgba + row flip reproduce the reference's own render."""
    g = json.load(open(GOLDEN / "png_sky_rows.json"))
    W, H = g["width"], g["height"]
    cam = oracle.camera_new(**{k: v for k, v in g["camera"].items() if k != "cite"})
    rows = sorted({s[1] for s in g["samples_x_row_r_g_b"]})
    # empty world: the fixture pixels see only sky, whatever random scene the reference rendered
    empty = oracle.Scene(np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.uint32), np.zeros(1, np.uint32), np.ones((1, 3)), np.zeros(1))
    img, _, _ = oracle.render(empty, cam, W, H, spp=8, seed=3, rows=(rows[0], rows[-1] + 1),
                              sampler=oracle.SAMPLER_DIRECT if sampler == "direct" else oracle.SAMPLER_REJECTION)
    worst = 0
    for x, y, r, gg, b in g["samples_x_row_r_g_b"]:
        worst = max(worst, int(np.abs(img[y, x, :3].astype(int) - np.array([r, gg, b])).max()))
        assert img[y, x, 3] == g["alpha"]
    assert worst <= g["tolerance_lsb"], f"sky pixels differ from the reference PNG by {worst} LSB"
    # orientation: PNG row 0 is the TOP (v ~ 1): lighter-blue-to-white gradient goes down the image
    assert img[rows[0], 0, 0] <= img[rows[-1], 0, 0]


@pytest.mark.parametrize("sampler", ["direct", "rejection"])
def test_png_big_spheres(oracle, sampler):
    """The non-random part of random_scene (ground + glass / Lambertian / Metal unit spheres, main.rs:62-64,93-99) as the
    reference itself rendered it: the oracle reproduces the metal sphere's sky-mirroring cap to a quarter of an LSB (391 block
    means: get_ray, Sphere::hit, Metal::scatter, the recursion, sky, to_rgba), the sky seen THROUGH the glass sphere
    (Dialectric::scatter: two refractions + Schlick) and the top of the diffuse sphere (Lambertian::scatter) within the few LSB by
    which the reference's random small spheres shade them (tests/golden/make_png_big_spheres_fixture.py)."""
    g = json.load(open(GOLDEN / "png_big_spheres.json"))
    s = g["scene"]
    sc = oracle.Scene(s["center"], s["radius"], [0, 1, 2, 3], s["mat_kind"], s["mat_albedo"], s["mat_param"])
    cam = oracle.camera_new(**{k: v for k, v in g["camera"].items() if k != "cite"})
    ys = [b[1] for r in g["regions"].values() for b in r["blocks_x_y_r_g_b"]]
    img, _, _ = oracle.render(sc, cam, g["width"], g["height"], spp=128, seed=7, rows=(min(ys) - 4, max(ys) + 5),
                              sampler=oracle.SAMPLER_DIRECT if sampler == "direct" else oracle.SAMPLER_REJECTION)
    for name, r in g["regions"].items():
        lo, hi = r["tolerance_lsb_lo_hi"]
        hw, hh = r["half_w"], r["half_h"]
        d = np.array([img[y - hh:y + hh + 1, x - hw:x + hw + 1, :3].reshape(-1, 3).astype(float).mean(0) - np.array(c) for x, y, *c in r["blocks_x_y_r_g_b"]])
        assert len(d) >= 30 and lo <= d.min() and d.max() <= hi, f"{name}: block means differ from the reference PNG by {d.min():.2f} .. {d.max():.2f} LSB"
        if "max_abs_mean" in r:
            assert np.abs(d.mean(0)).max() <= r["max_abs_mean"], f"{name}: biased against the reference PNG by {d.mean(0)} LSB"


def test_png_corner_pixels_exact(oracle):
    """(0,0),(600,0),(1199,0) -> [220,235,255]; (0,39) -> [221,235,255]."""
    cam = final_camera(oracle, 1.5)
    empty = oracle.Scene(np.zeros((0, 3)), np.zeros(0), np.zeros(0, np.uint32), np.zeros(1, np.uint32), np.ones((1, 3)), np.zeros(1))
    img, _, _ = oracle.render(empty, cam, 1200, 800, spp=16, seed=1, rows=(0, 40))
    for (x, y), want in {(0, 0): [220, 235, 255], (600, 0): [220, 235, 255], (1199, 0): [220, 235, 255], (0, 39): [221, 235, 255]}.items():
        assert list(img[y, x, :3]) == want


# --------------------------------------------------------------------------------------------- (2) analytic KATs
def test_sphere_hit_axis_ray(oracle):
    # axis-aligned ray vs unit sphere at the origin: t = |o| - 1, outward normal, front face
    r = oracle.sphere_hit_batch([[0, 0, 0]], 1.0, [[0, 0, -5]], [[0, 0, 1]], 1e-4, np.inf)
    assert r["hit"][0] == 1 and r["t"][0] == pytest.approx(4.0, abs=1e-15)
    assert np.allclose(r["p"][0], [0, 0, -1]) and np.allclose(r["normal"][0], [0, 0, -1]) and r["front_face"][0] == 1


def test_sphere_hit_unnormalised_direction(oracle):
    # t is in units of |dir| (Appendix C.3): dir = (0,0,2) halves t
    r = oracle.sphere_hit_batch([[0, 0, 0]], 1.0, [[0, 0, -5]], [[0, 0, 2]], 1e-4, np.inf)
    assert r["t"][0] == pytest.approx(2.0, abs=1e-15) and np.allclose(r["p"][0], [0, 0, -1])


def test_sphere_hit_from_inside(oracle):
    # origin inside: first root < t_min -> second root, back face, normal flipped against the ray
    r = oracle.sphere_hit_batch([[0, 0, 0]], 1.0, [[0, 0, 0]], [[1, 0, 0]], 1e-4, np.inf)
    assert r["hit"][0] == 1 and r["t"][0] == pytest.approx(1.0) and r["front_face"][0] == 0
    assert np.allclose(r["normal"][0], [-1, 0, 0])


def test_sphere_hit_tangent_and_miss(oracle):
    r = oracle.sphere_hit_batch([[0, 0, 0], [0, 0, 0]], 1.0, [[1, 0, -5], [1.0000001, 0, -5]], [[0, 0, 1], [0, 0, 1]], 1e-4, np.inf)
    assert list(r["hit"]) == [1, 0] and r["t"][0] == pytest.approx(5.0)      # disc == 0 is a hit (sphere.rs:25 is `< 0.0`)


def test_sphere_hit_tmax_inclusive(oracle):
    # root == t_max is ACCEPTED (sphere.rs:29: `t_max < root` rejects) ; root just above is rejected
    r = oracle.sphere_hit_batch([[0, 0, 0]] * 2, 1.0, [[0, 0, -5]] * 2, [[0, 0, 1]] * 2, 1e-4, [4.0, np.nextafter(4.0, 0)])
    assert list(r["hit"]) == [1, 0]


def test_sphere_hit_negative_radius(oracle):
    # negative radius is legal (sphere.rs:45-51) and flips the outward normal -> hollow glass
    r = oracle.sphere_hit_batch([[0, 0, 0]], -1.0, [[0, 0, -5]], [[0, 0, 1]], 1e-4, np.inf)
    assert r["hit"][0] == 1 and r["t"][0] == pytest.approx(4.0) and r["front_face"][0] == 0
    assert np.allclose(r["normal"][0], [0, 0, -1])


def test_world_hit_tie_later_index_wins(oracle):
    # two identical spheres: exact tie -> the LATER list entry wins (Appendix C.2)
    sc = oracle.Scene([[0, 0, 0], [0, 0, 0], [0, 0, 10]], [1, 1, 1], [0, 0, 0], [0], [[1, 1, 1]], [0])
    r = oracle.world_hit_batch(sc, [[0, 0, -5]], [[0, 0, 1]])
    assert r["hit"][0] == 1 and r["index"][0] == 1 and r["t"][0] == pytest.approx(4.0)


def test_reflect_refract_known_answers(oracle):
    v = np.array([[1.0, -1.0, 0.0]]); n = np.array([[0.0, 1.0, 0.0]])
    assert np.allclose(oracle.reflect_batch(v, n), [[1, 1, 0]])
    rv = oracle.reflect_batch(np.random.default_rng(0).normal(size=(100, 3)), np.tile(n, (100, 1)))
    assert np.allclose(np.linalg.norm(rv, axis=1), np.linalg.norm(np.random.default_rng(0).normal(size=(100, 3)), axis=1))
    # normal incidence: direction unchanged
    assert np.allclose(oracle.refract_batch([[0, -1, 0]], n, 1 / 1.5), [[0, -1, 0]])
    # Snell: eta * sin(theta_i) = sin(theta_t)
    th = np.radians(40.0)
    out = oracle.refract_batch([[np.sin(th), -np.cos(th), 0]], n, 1 / 1.5)[0]
    assert np.linalg.norm(out) == pytest.approx(1.0) and out[0] == pytest.approx(np.sin(th) / 1.5)


def test_schlick_and_dielectric_branches(oracle):
    L = oracle.lib()
    assert L.o_reflectance(1.0, 1 / 1.5) == pytest.approx(0.04)           # r0 = ((1-eta)/(1+eta))^2
    assert L.o_reflectance(0.0, 1 / 1.5) == pytest.approx(1.0)
    n = [[0, 1, 0]]
    # head-on: R = 0.04; xi = 0.03 -> reflect (R <= xi false), xi = 0.04.. -> refract (`<=`, materials.rs:96)
    kw = dict(kind=[2], albedo=[[0, 0, 0]], param=[1.5], r_orig=[[0, 5, 0]], r_dir=[[0, -2, 0]], p=[[0, 0, 0]], normal=n, front_face=[1])
    refl = oracle.scatter_batch(sample=[[0.03, 0, 0]], **kw)
    refr = oracle.scatter_batch(sample=[[0.05, 0, 0]], **kw)
    assert np.allclose(refl["dir"], [[0, 1, 0]]) and np.allclose(refr["dir"], [[0, -1, 0]])
    assert np.allclose(refl["attenuation"], 1) and refl["some"][0] == 1
    # total internal reflection from inside (front_face = 0 -> ratio = ir): xi ignored
    th = np.radians(60.0)
    tir = oracle.scatter_batch(kind=[2], albedo=[[0, 0, 0]], param=[1.5], r_orig=[[0, 0, 0]], r_dir=[[np.sin(th), -np.cos(th), 0]],
                               p=[[0, 0, 0]], normal=n, front_face=[0], sample=[[0.999, 0, 0]])
    assert np.allclose(tir["dir"], [[np.sin(th), np.cos(th), 0]])


def test_lambertian_and_metal_branches(oracle):
    n = [[0, 1, 0]]
    # Lambertian: dir = n + unit(sample); degenerate guard when sample = -n (vec3.rs:111-114)
    r = oracle.scatter_batch([0, 0], [[.1, .2, .3]] * 2, [0, 0], [[0, 1, 0]] * 2, [[0, -1, 0]] * 2, [[0, 0, 0]] * 2, n * 2, [1, 1],
                             [[0.5, 0, 0], [0, -0.25, 0]])
    assert np.allclose(r["dir"], [[1, 1, 0], [0, 1, 0]]) and np.allclose(r["attenuation"], [[.1, .2, .3]] * 2) and list(r["some"]) == [1, 1]
    # Metal: fuzz 0 mirror; absorbed when dir.n <= 0; fuzz is NOT clamped
    m = oracle.scatter_batch([1, 1], [[.7, .6, .5]] * 2, [0.0, 3.0], [[0, 1, 0]] * 2, [[2, -2, 0]] * 2, [[0, 0, 0]] * 2, n * 2, [1, 1],
                             [[0.3, 0.3, 0.3], [0, -0.9, 0]])
    s = np.sqrt(0.5)
    assert np.allclose(m["dir"][0], [s, s, 0]) and m["some"][0] == 1
    assert m["some"][1] == 0 and np.allclose(m["dir"][1], [s, s - 2.7, 0])


def test_to_rgba_known_answers(oracle):
    spp = 100
    out = oracle.to_rgba_batch([[spp, 0.25 * spp, 0], [4 * spp, -1, np.nan], [0.999 ** 2 * spp, 1e-12, 0.5 * spp]], 255, spp)
    assert list(out[0]) == [255, 128, 0, 255]                 # sqrt(1)=1 -> clamp .999 -> 255; sqrt(.25)=.5 -> 128
    assert list(out[1]) == [255, 0, 0, 255]                   # >1 clamps; sqrt(neg)=NaN -> 0; NaN -> 0
    assert out[2][2] == int(256 * np.sqrt(0.5))


def test_camera_basis_and_rays(oracle):
    cam = final_camera(oracle, 1.5)
    u, v, w = cam.u.np(), cam.v.np(), cam.w.np()
    assert np.allclose([u @ v, u @ w, v @ w], 0, atol=1e-15) and np.allclose([u @ u, v @ v, w @ w], 1)
    assert cam.lens_radius == 0.05
    # centre ray with no lens offset points at look_at and reaches it at t = 1 (focus plane at focus_dist = 10)
    r = oracle.get_ray_batch(cam, [0.5], [0.5], [[0, 0]])
    assert np.allclose(r["orig"][0], [13, 2, 3])
    d = r["dir"][0]
    assert np.linalg.norm(d) == pytest.approx(10.0) and np.allclose(np.cross(d, [-13, -2, -3]), 0, atol=1e-12)
    # lens offset moves origin and direction oppositely: origin + dir is invariant (thin lens)
    r2 = oracle.get_ray_batch(cam, [0.3], [0.8], [[0.6, -0.2]])
    r3 = oracle.get_ray_batch(cam, [0.3], [0.8], [[0, 0]])
    assert np.allclose(r2["orig"] + r2["dir"], r3["orig"] + r3["dir"])


def test_ray_color_depth_and_sky(oracle, final_scene):
    _, sc = final_scene
    up = oracle.ray_color_batch(sc, [[0, 5, 0]], [[0, 3, 0]], [7], [0], seed=1)
    assert np.allclose(up["color"][0], [0.5, 0.7, 1.0]) and up["rays"][0] == 1          # t = 1 -> (0.5,0.7,1.0)
    zero = oracle.ray_color_batch(sc, [[0, 5, 0]], [[0, -1, 0]], [7], [0], seed=1, max_depth=0)
    assert np.allclose(zero["color"], 0) and zero["rays"][0] == 0                        # main.rs:40-42
    one = oracle.ray_color_batch(sc, [[0, 5, 0]], [[0, -1, 0]], [7], [0], seed=1, max_depth=1)
    assert np.allclose(one["color"], 0) and one["rays"][0] == 1                          # hit, scatter, then depth 0 -> black


# --------------------------------------------------------------------------------------------- (3) Philox KATs
PHILOX_KAT = [   # Random123 kat_vectors, philox4x32 10 rounds
    ([0, 0, 0, 0], [0, 0], [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]),
    ([0xffffffff] * 4, [0xffffffff] * 2, [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]),
    ([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0], [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]),
]


@pytest.mark.parametrize("ctr,key,want", PHILOX_KAT)
def test_philox_known_answers(oracle, ctr, key, want):
    assert [int(x) for x in oracle.philox4x32_10(ctr, key)] == want
    assert npr.philox4x32_10(ctr, key) == want


def test_direct_uniforms_are_24_bit(oracle):
    u = oracle.direct_uniforms(0x1234567800000001, 5, 6, 7)
    words = npr.philox4x32_10([5, 6, 7, 0], [1, 0x12345678])
    assert np.array_equal(u, [(w >> 8) / 2.0 ** 24 for w in words])
    assert np.array_equal(u.astype(np.float32).astype(np.float64), u)        # exactly representable in f32


# --------------------------------------------------------------------------------------------- (4) numpy restatement
def _rand_rays(rng, n):
    o = rng.uniform(-12, 12, (n, 3)) * [1, 0.2, 1] + [0, 1.5, 0]
    d = rng.normal(size=(n, 3)) * rng.uniform(0.2, 3.0, (n, 1))
    return o, d


def test_sphere_hit_vs_numpy(oracle):
    rng = np.random.default_rng(1)
    n = 20000
    c = rng.uniform(-5, 5, (n, 3)); r = rng.uniform(0.1, 3, n) * rng.choice([1, 1, 1, -1], n)
    o, d = _rand_rays(rng, n)
    d = c - o + rng.normal(size=(n, 3)) * np.abs(r)[:, None] * 0.7        # aim near the sphere so ~half hit
    a = oracle.sphere_hit_batch(c, r, o, d, 1e-4, np.inf)
    hit, t, p, nrm, ff = npr.sphere_hit(c, r, o, d, 1e-4, np.inf)
    assert 0.2 < hit.mean() < 0.9
    assert np.array_equal(a["hit"].astype(bool), hit)
    m = hit
    assert np.array_equal(a["front_face"][m].astype(bool), ff[m])
    assert rel_err(a["t"][m], t[m]).max() < 1e-12 and np.abs(a["p"][m] - p[m]).max() < 1e-11 and np.abs(a["normal"][m] - nrm[m]).max() < 1e-11


def test_world_hit_vs_numpy(oracle, final_scene):
    arrays, sc = final_scene
    rng = np.random.default_rng(2)
    o, d = _rand_rays(rng, 3000)
    a = oracle.world_hit_batch(sc, o, d)
    idx, t = npr.world_hit(arrays["center"], arrays["radius"], o, d, 1e-4)
    assert np.array_equal(a["index"], idx)
    m = idx >= 0
    assert 0.3 < m.mean() < 0.99 and rel_err(a["t"][m], t[m]).max() < 1e-12


def test_scatter_vs_numpy(oracle):
    rng = np.random.default_rng(3)
    n = 6000
    nrm = npr.unit(rng.normal(size=(n, 3)))
    rd = rng.normal(size=(n, 3)) * rng.uniform(0.1, 4, (n, 1))
    rd = np.where((npr.dot(rd, nrm) > 0)[:, None], -rd, rd)               # rays arrive against the normal
    smp = rng.uniform(-1, 1, (n, 3)); smp *= (rng.uniform(0, 1, (n, 1)) ** (1 / 3)) / np.linalg.norm(smp, axis=1, keepdims=True)
    p = rng.uniform(-3, 3, (n, 3)); alb = rng.uniform(0, 1, (n, 3)); ff = rng.integers(0, 2, n)
    lam = oracle.scatter_batch(np.zeros(n, int), alb, np.zeros(n), p, rd, p, nrm, ff, smp)
    assert np.abs(lam["dir"] - npr.scatter_lambertian(nrm, smp)).max() < 1e-13 and np.array_equal(lam["attenuation"], alb) and lam["some"].all()
    fuzz = rng.uniform(0, 1.5, n)
    met = oracle.scatter_batch(np.ones(n, int), alb, fuzz, p, rd, p, nrm, ff, smp)
    dm, some = npr.scatter_metal(rd, nrm, fuzz, smp)
    assert np.array_equal(met["some"].astype(bool), some) and np.abs(met["dir"] - dm).max() < 1e-13 and 0.02 < (~some).mean() < 0.6
    ir = rng.uniform(1.1, 2.4, n); xi = rng.uniform(0, 1, n)
    die = oracle.scatter_batch(np.full(n, 2), alb, ir, p, rd, p, nrm, ff, np.stack([xi, xi * 0, xi * 0], 1))
    assert np.abs(die["dir"] - npr.scatter_dielectric(rd, nrm, ff.astype(bool), ir, xi)).max() < 1e-12
    assert np.allclose(die["attenuation"], 1) and np.array_equal(die["orig"], p)


def test_camera_and_rgba_vs_numpy(oracle):
    args = ((13, 2, 3), (0, 0, 0), (0, 1, 0), 20.0, 16 / 9, 0.1, 10.0)
    cam = oracle.camera_new(*args); ref = npr.camera_new(*args)
    for k, name in [("origin", "origin"), ("lower_left_corner", "llc"), ("horizontal", "horizontal"), ("vertical", "vertical"), ("u", "u"), ("v", "v"), ("w", "w")]:
        assert np.abs(getattr(cam, k).np() - ref[name]).max() < 1e-14
    rng = np.random.default_rng(4)
    s, t, disk = rng.uniform(0, 1, 500), rng.uniform(0, 1, 500), rng.uniform(-0.7, 0.7, (500, 2))
    a = oracle.get_ray_batch(cam, s, t, disk); o, d = npr.get_ray(ref, s, t, disk)
    assert np.abs(a["orig"] - o).max() < 1e-13 and np.abs(a["dir"] - d).max() < 1e-13
    col = rng.uniform(0, 130, (5000, 3))
    assert np.array_equal(oracle.to_rgba_batch(col, 200, 100), npr.to_rgba(col, 200, 100))


def test_ray_color_recursion_vs_manual_unroll(oracle, final_scene):
    """ray_color's recursion (main.rs:49) equals a hand-unrolled throughput product built from the unit functions."""
    arrays, sc = final_scene
    cam = final_camera(oracle, 16 / 9)
    rng = np.random.default_rng(5)
    n = 300
    r = oracle.get_ray_batch(cam, rng.uniform(0, 1, n), rng.uniform(0, 0.6, n), np.zeros((n, 2)))
    pix = np.arange(n, dtype=np.uint32); smp = np.full(n, 3, np.uint32)
    got = oracle.ray_color_batch(sc, r["orig"], r["dir"], pix, smp, seed=9, max_depth=50)
    for i in range(0, n, 7):
        o, d = r["orig"][i:i + 1], r["dir"][i:i + 1]
        thr, col, rays = np.ones(3), np.zeros(3), 0
        for bounce in range(50):
            h = oracle.world_hit_batch(sc, o, d); rays += 1
            if not h["hit"][0]:
                col = thr * npr.sky(d)[0]; break
            k = h["index"][0]; u = oracle.direct_uniforms(9, int(pix[i]), 3, bounce + 1)
            kind = int(arrays["mat_kind"][k])
            L = oracle.lib()
            smpv = (L.o_direct_unit_vector(u[0], u[1]).np() if kind == 0 else L.o_direct_in_unit_sphere(u[0], u[1], u[2]).np() if kind == 1
                    else np.array([u[0], 0, 0]))
            s = oracle.scatter_batch([kind], arrays["mat_albedo"][k:k + 1], arrays["mat_param"][k:k + 1], o, d, h["p"], h["normal"],
                                     h["front_face"], [smpv])
            if not s["some"][0]:
                break
            thr = thr * s["attenuation"][0]; o, d = s["orig"], s["dir"]
        assert rays == got["rays"][i]
        assert np.abs(col - got["color"][i]).max() < 1e-12


# --------------------------------------------------------------------------------------------- (5) sampler equivalence
def _moments(v):
    r = np.linalg.norm(v, axis=1)
    return r, v.mean(axis=0), (v ** 2).mean(axis=0)


def test_inversion_samplers_match_rejection_samplers(oracle):
    """Distribution identity (Appendix B consequence 2): uniform-in-disk / on-sphere / in-ball."""
    n = 200_000
    L = oracle.lib()
    U = np.array([oracle.direct_uniforms(11, i, 0, 1) for i in range(20000)])
    # rejection draws (reference loops)
    disk_r = oracle.rejection_samples(5, n, 0); ball_r = oracle.rejection_samples(6, n, 1); unit_r = oracle.rejection_samples(7, n, 2)
    assert (disk_r[:, 2] == 0).all() and (np.linalg.norm(disk_r, axis=1) < 1).all() and (np.linalg.norm(ball_r, axis=1) < 1).all()
    assert np.allclose(np.linalg.norm(unit_r, axis=1), 1)
    # inversion draws from the Philox uniforms
    import ctypes as C
    dx, dy = C.c_double(), C.c_double()
    disk_d, unit_d, ball_d = [], [], []
    for u in U:
        L.o_direct_disk(u[2], u[3], C.byref(dx), C.byref(dy)); disk_d.append([dx.value, dy.value, 0.0])
        unit_d.append(L.o_direct_unit_vector(u[0], u[1]).np()); ball_d.append(L.o_direct_in_unit_sphere(u[0], u[1], u[2]).np())
    disk_d, unit_d, ball_d = map(np.array, (disk_d, unit_d, ball_d))
    tol = 4.0 / np.sqrt(len(U))                                   # ~4 sigma of the smaller sample
    # disk: E r^2 = 1/2, E x^2 = 1/4; ball: E r^3... use cdf checks: P(r<a) = a^2 (disk), a^3 (ball)
    for a in (0.3, 0.6, 0.9):
        assert abs((np.linalg.norm(disk_d, axis=1) < a).mean() - a ** 2) < tol and abs((np.linalg.norm(disk_r, axis=1) < a).mean() - a ** 2) < tol
        assert abs((np.linalg.norm(ball_d, axis=1) < a).mean() - a ** 3) < tol and abs((np.linalg.norm(ball_r, axis=1) < a).mean() - a ** 3) < tol
    for d_, r_ in ((disk_d, disk_r), (unit_d, unit_r), (ball_d, ball_r)):
        _, m1d, m2d = _moments(d_); _, m1r, m2r = _moments(r_)
        assert np.abs(m1d - m1r).max() < tol and np.abs(m2d - m2r).max() < tol
    # z of a uniform direction is uniform on [-1,1] (Archimedes) in both
    for z in (unit_d[:, 2], unit_r[:, 2]):
        assert abs((z < 0.25).mean() - 0.625) < tol
    assert np.allclose(np.linalg.norm(unit_d, axis=1), 1, atol=1e-12)


def test_oracle_direct_and_rejection_renders_agree_statistically(oracle, final_scene):
    """Same scene, the two samplers: difference is Monte-Carlo noise only (no bias)."""
    _, sc = final_scene
    cam = final_camera(oracle, 16 / 9)
    a, acc_a, _ = oracle.render(sc, cam, 160, 90, spp=64, seed=1, sampler=oracle.SAMPLER_DIRECT, want_accum=True)
    b, acc_b, _ = oracle.render(sc, cam, 160, 90, spp=64, seed=2, sampler=oracle.SAMPLER_REJECTION, want_accum=True)
    c, acc_c, _ = oracle.render(sc, cam, 160, 90, spp=64, seed=3, sampler=oracle.SAMPLER_REJECTION, want_accum=True)
    rmse = lambda x, y: float(np.sqrt(((x[..., :3].astype(float) - y[..., :3].astype(float)) ** 2).mean()))
    floor = rmse(b, c)
    assert rmse(a, b) < 1.25 * floor and rmse(a, c) < 1.25 * floor
    mean_a, mean_b = acc_a.mean(axis=(0, 1)) / 64, acc_b.mean(axis=(0, 1)) / 64
    assert np.abs(mean_a - mean_b).max() < 0.01          # radiance units; noise of the mean ~ 2e-3
