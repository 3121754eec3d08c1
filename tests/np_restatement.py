"""Independent numpy restatement of the reference's hot-path functions, written from the Rust sources
without looking at oracle/rtiow_oracle.c, to guard against a shared misreading (SURVEY §7 hard part 6).
Vectorised over the leading axis; f64 throughout.  Test infrastructure only.
"""
import numpy as np


def dot(a, b):
    return a[..., 0] * b[..., 0] + a[..., 1] * b[..., 1] + a[..., 2] * b[..., 2]          # vec3.rs:95-97


def length(a):
    return np.sqrt(dot(a, a))                                                              # vec3.rs:83-89


def unit(a):
    return a * (1.0 / length(a))[..., None]                                                # vec3.rs:107-109, 371-376


def reflect(v, n):
    return v - (2.0 * dot(v, n))[..., None] * n                                            # vec3.rs:116-118


def refract(uv, n, eta):
    cos_theta = np.minimum(1.0, -dot(uv, n))                                               # vec3.rs:121
    perp = eta[..., None] * (uv + n * cos_theta[..., None])                                # vec3.rs:122
    par = -np.sqrt(np.abs(1.0 - dot(perp, perp)))[..., None] * n                           # vec3.rs:123
    return perp + par


def sphere_hit(center, radius, orig, direction, t_min, t_max):
    """sphere.rs:16-41 + mod.rs:20-30 -> hit, t, p, normal, front_face"""
    oc = orig - center
    a = dot(direction, direction)
    half_b = dot(oc, direction)
    c = dot(oc, oc) - radius * radius
    disc = half_b * half_b - a * c
    ok = disc >= 0.0
    sq = np.sqrt(np.where(ok, disc, 0.0))
    r1 = (-half_b - sq) / a
    r2 = (-half_b + sq) / a
    bad1 = (r1 < t_min) | (t_max < r1)
    bad2 = (r2 < t_min) | (t_max < r2)
    root = np.where(bad1, r2, r1)
    hit = ok & ~(bad1 & bad2)
    p = orig + root[..., None] * direction
    outward = (p - center) * (1.0 / radius)[..., None]
    ff = dot(direction, outward) < 0.0
    normal = np.where(ff[..., None], outward, -outward)
    return hit, root, p, normal, ff


def world_hit(centers, radii, orig, direction, t_min, t_max=np.inf):
    """mod.rs:56-69 for a batch of rays: sequential over spheres, shrinking t_max, ties -> later index"""
    n = len(orig)
    closest = np.full(n, t_max)
    idx = np.full(n, -1)
    for i in range(len(radii)):
        h, t, _, _, _ = sphere_hit(centers[i][None, :], np.full(n, radii[i]), orig, direction, t_min, closest)
        closest = np.where(h, t, closest)
        idx = np.where(h, i, idx)
    return idx, closest


def reflectance(cosine, ref_idx):
    r0 = ((1.0 - ref_idx) / (1.0 + ref_idx)) ** 2                                          # materials.rs:79
    return r0 + (1.0 - r0) * (1.0 - cosine) ** 5                                           # materials.rs:80


def scatter_lambertian(normal, sample):
    d = normal + unit(sample)                                                              # materials.rs:23
    near = (np.abs(d) < 1e-8).all(axis=-1)                                                 # vec3.rs:111-114
    return np.where(near[..., None], normal, d)


def scatter_metal(r_dir, normal, fuzz, sample):
    refl = unit(reflect(r_dir, normal))                                                    # materials.rs:51
    d = refl + fuzz[..., None] * sample                                                    # materials.rs:53
    return d, dot(d, normal) > 0.0                                                         # materials.rs:55-57


def scatter_dielectric(r_dir, normal, front_face, ir, xi):
    ratio = np.where(front_face, 1.0 / ir, ir)                                             # materials.rs:84-87
    ud = unit(r_dir)
    cos_theta = np.minimum(1.0, -dot(ud, normal))
    sin_theta = np.sqrt(1.0 - cos_theta * cos_theta)
    can_refract = ratio * sin_theta <= 1.0
    do_refract = can_refract & (reflectance(cos_theta, ratio) <= xi)                       # materials.rs:96
    return np.where(do_refract[..., None], refract(ud, normal, ratio), reflect(ud, normal))


def camera_new(look_from, look_at, v_up, v_fov, aspect, aperture, focus_dist):
    look_from, look_at, v_up = (np.asarray(x, float) for x in (look_from, look_at, v_up))
    theta = np.radians(v_fov)
    vh = 2.0 * np.tan(theta / 2.0)
    vw = aspect * vh
    w = unit(look_from - look_at)
    u = unit(np.cross(v_up, w))
    v = np.cross(w, u)
    hor = focus_dist * vw * u
    ver = focus_dist * vh * v
    llc = look_from - hor / 2.0 - ver / 2.0 - focus_dist * w
    return dict(origin=look_from, llc=llc, horizontal=hor, vertical=ver, u=u, v=v, w=w, lens_radius=aperture / 2.0)


def get_ray(cam, s, t, disk):
    rd = cam["lens_radius"] * disk                                                         # camera.rs:48
    offset = cam["u"] * rd[..., 0:1] + cam["v"] * rd[..., 1:2]
    orig = cam["origin"] + offset
    direction = cam["llc"] + s[..., None] * cam["horizontal"] + t[..., None] * cam["vertical"] - cam["origin"] - offset
    return orig, direction


def sky(direction):
    t = 0.5 * (unit(direction)[..., 1] + 1.0)                                              # main.rs:54-56
    return (1.0 - t)[..., None] * np.array([1.0, 1.0, 1.0]) + t[..., None] * np.array([0.5, 0.7, 1.0])


def to_rgba(color_sum, alpha, spp):
    c = np.sqrt(color_sum * (1.0 / spp))                                                   # vec3.rs:410-413
    q = (256.0 * np.clip(c, 0.0, 0.999))
    q = np.where(np.isnan(q), 0.0, q).astype(np.uint8)                                     # `as u8`: truncate, NaN -> 0
    return np.concatenate([q, np.full(q.shape[:-1] + (1,), alpha, np.uint8)], axis=-1)


def philox4x32_10(ctr, key):
    """Random123 Philox4x32-10, pure Python ints (independent of the C and CUDA implementations)."""
    c = [int(x) for x in ctr]
    k = [int(x) for x in key]
    M0, M1, W0, W1, MASK = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85, 0xFFFFFFFF
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [(p1 >> 32) ^ c[1] ^ k[0], p1 & MASK, (p0 >> 32) ^ c[3] ^ k[1], p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c
