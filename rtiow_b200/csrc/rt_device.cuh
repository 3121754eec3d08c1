// rt_device.cuh — device-side building blocks of the rtiow render path (sm_100a).
//
// Every function here is the B200 form of one reference function; the citation is
// /root/reference/src/<file>:<lines>.  All functions are templated on real_t: float is the
// product path, double restates the reference's f64 arithmetic for parity triage and is also what
// the float renderer uses for ill-conditioned (large-radius) spheres.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rt {

// ------------------------------------------------------------------------------------------------
// Vec3 (vec3.rs:4-9) and the operators the hot path uses (vec3.rs:137-397)
// ------------------------------------------------------------------------------------------------
template <typename T> struct V3 { T x, y, z; };

template <typename T> __host__ __device__ __forceinline__ V3<T> mk(T x, T y, T z) { V3<T> r; r.x = x; r.y = y; r.z = z; return r; }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator+(V3<T> a, V3<T> b) { return mk<T>(a.x + b.x, a.y + b.y, a.z + b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator-(V3<T> a, V3<T> b) { return mk<T>(a.x - b.x, a.y - b.y, a.z - b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, T s) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(T s, V3<T> a) { return mk<T>(a.x * s, a.y * s, a.z * s); }
template <typename T> __host__ __device__ __forceinline__ V3<T> operator*(V3<T> a, V3<T> b) { return mk<T>(a.x * b.x, a.y * b.y, a.z * b.z); }
template <typename T> __host__ __device__ __forceinline__ V3<T> neg(V3<T> a) { return mk<T>(-a.x, -a.y, -a.z); }
template <typename T> __host__ __device__ __forceinline__ T dot(V3<T> a, V3<T> b) { return a.x * b.x + a.y * b.y + a.z * b.z; }          // vec3.rs:95-97
template <typename T> __host__ __device__ __forceinline__ T length_squared(V3<T> a) { return dot(a, a); }                                 // vec3.rs:87-89

__device__ __forceinline__ float rsqrt_t(float x) { return rsqrtf(x); }
__device__ __forceinline__ double rsqrt_t(double x) { return 1.0 / sqrt(x); }
__device__ __forceinline__ float sqrt_t(float x) { return sqrtf(x); }
__device__ __forceinline__ double sqrt_t(double x) { return sqrt(x); }
__device__ __forceinline__ float cbrt_t(float x) { return cbrtf(x); }
__device__ __forceinline__ double cbrt_t(double x) { return cbrt(x); }
__device__ __forceinline__ void sincospi_t(float x, float* s, float* c) { sincospif(x, s, c); }
__device__ __forceinline__ void sincospi_t(double x, double* s, double* c) { sincospi(x, s, c); }
__device__ __forceinline__ float min_t(float a, float b) { return fminf(a, b); }
__device__ __forceinline__ double min_t(double a, double b) { return fmin(a, b); }
__device__ __forceinline__ float max_t(float a, float b) { return fmaxf(a, b); }
__device__ __forceinline__ double max_t(double a, double b) { return fmax(a, b); }
__device__ __forceinline__ float abs_t(float a) { return fabsf(a); }
__device__ __forceinline__ double abs_t(double a) { return fabs(a); }

template <typename T> __device__ __forceinline__ T length(V3<T> a) { return sqrt_t(length_squared(a)); }                                   // vec3.rs:83-85
// unit_vector (vec3.rs:107-109): self / length, and Div<f64> is `self * (1.0/scalar)` (vec3.rs:371-376)
template <typename T> __device__ __forceinline__ V3<T> unit_vector(V3<T> a) { return a * (T(1) / length(a)); }
template <typename T> __device__ __forceinline__ V3<T> cross(V3<T> a, V3<T> b)                                                          // vec3.rs:99-105
{
    return mk<T>(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
template <typename T> __device__ __forceinline__ bool is_near_zero(V3<T> a)                                                             // vec3.rs:111-114
{
    const T s = T(1e-8);
    return abs_t(a.x) < s && abs_t(a.y) < s && abs_t(a.z) < s;
}
template <typename T> __device__ __forceinline__ V3<T> reflect(V3<T> v, V3<T> n) { return v - n * (T(2) * dot(v, n)); }                   // vec3.rs:116-118
template <typename T> __device__ __forceinline__ V3<T> refract(V3<T> uv, V3<T> n, T etai_over_etat)                                      // vec3.rs:120-125
{
    T cos_theta = min_t(T(1), -dot(uv, n));
    V3<T> r_out_perp = (uv + n * cos_theta) * etai_over_etat;
    V3<T> r_out_parallel = n * (-sqrt_t(abs_t(T(1) - length_squared(r_out_perp))));
    return r_out_perp + r_out_parallel;
}

// Color::to_rgba (vec3.rs:404-420).  `sum` is the per-pixel radiance sum; returns packed RGBA
// (R in the low byte).  Rust `as u8` truncates, saturates and maps NaN to 0.
template <typename T> __device__ __forceinline__ uint32_t quantise_channel(T sum, T scale)
{
    T c = sqrt_t(scale * sum);
    if (c != c) return 0u;                          // NaN.clamp() stays NaN; NaN as u8 == 0
    c = c < T(0) ? T(0) : (c > T(0.999) ? T(0.999) : c);   // f64::clamp(0.0, 0.999)
    T v = T(256) * c;
    return v >= T(255) ? 255u : (uint32_t)(int)v;
}
template <typename T> __device__ __forceinline__ uint32_t to_rgba(V3<T> sum, uint32_t alpha, uint64_t spp)
{
    T scale = T(1) / T(spp);
    return quantise_channel(sum.x, scale) | (quantise_channel(sum.y, scale) << 8) | (quantise_channel(sum.z, scale) << 16) | (alpha << 24);
}

// ------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11): the counter-based stream that replaces thread_rng
// (vec3.rs:22,27,61; materials.rs:95; main.rs:128).  One block per EVENT of a path, keyed
// (seed; pixel, sample, bounce) so the image does not depend on thread, block or GPU count.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// The ten round keys of a seed, expanded once (on the host for the render kernel: they arrive as kernel parameters, i.e.
// as constant-bank operands of the XORs, instead of 18 uniform-datapath additions per block).
struct PhiloxKey { uint32_t k[20]; };
__host__ __device__ __forceinline__ PhiloxKey philox_key(uint64_t seed)
{
    PhiloxKey key; uint32_t k0 = (uint32_t)seed, k1 = (uint32_t)(seed >> 32);
    for (int r = 0; r < 10; ++r) { key.k[2 * r] = k0; key.k[2 * r + 1] = k1; k0 += 0x9E3779B9u; k1 += 0xBB67AE85u; }
    return key;
}
__host__ __device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, const PhiloxKey& key, uint32_t out[4])
{
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c0;
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c2;
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ key.k[2 * r];
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ key.k[2 * r + 1];
        c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// 24-bit uniform in [0,1): exactly representable in float and double, so the float renderer, the
// double renderer and the CPU oracle consume identical random numbers.
template <typename T> __host__ __device__ __forceinline__ T u01(uint32_t x) { return T(x >> 8) * T(1.0 / 16777216.0); }

template <typename T> struct Uniform4 { T u0, u1, u2, u3; };
template <typename T> __device__ __forceinline__ Uniform4<T> event_uniforms(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t bounce)
{
    uint32_t o[4];
    philox4x32_10(pixel, sample, bounce, 0u, (uint32_t)seed, (uint32_t)(seed >> 32), o);
    Uniform4<T> u; u.u0 = u01<T>(o[0]); u.u1 = u01<T>(o[1]); u.u2 = u01<T>(o[2]); u.u3 = u01<T>(o[3]);
    return u;
}
template <typename T> __device__ __forceinline__ Uniform4<T> event_uniforms(const PhiloxKey& key, uint32_t pixel, uint32_t sample, uint32_t bounce)
{
    uint32_t o[4];
    philox4x32_10(pixel, sample, bounce, 0u, key, o);
    Uniform4<T> u; u.u0 = u01<T>(o[0]); u.u1 = u01<T>(o[1]); u.u2 = u01<T>(o[2]); u.u3 = u01<T>(o[3]);
    return u;
}

// Inversion sampling of the distributions the reference draws by rejection:
//  random_in_unit_disk   (vec3.rs:59-68): uniform on the unit disk
//  random_unit_vector    (vec3.rs:47-49): uniform on the unit sphere
//  random_in_unit_sphere (vec3.rs:37-45): uniform in the unit ball
template <typename T> __device__ __forceinline__ void direct_disk(T u2, T u3, T* x, T* y)
{
    T s, c; sincospi_t(T(2) * u3, &s, &c);
    T r = sqrt_t(u2);
    *x = r * c; *y = r * s;
}
template <typename T> __device__ __forceinline__ V3<T> direct_unit_vector(T u0, T u1)
{
    T z = T(1) - T(2) * u0;
    T rxy = sqrt_t(max_t(T(0), T(1) - z * z));
    T s, c; sincospi_t(T(2) * u1, &s, &c);
    return mk<T>(rxy * c, rxy * s, z);
}
template <typename T> __device__ __forceinline__ V3<T> direct_in_unit_sphere(T u0, T u1, T u2) { return direct_unit_vector(u0, u1) * cbrt_t(u2); }

// ------------------------------------------------------------------------------------------------
// Camera (camera.rs:4-13, 47-54)
// ------------------------------------------------------------------------------------------------
template <typename T> struct CameraT {
    V3<T> origin, llc_minus_origin, horizontal, vertical, u, v;
    T lens_radius;
};
// Camera::get_ray (camera.rs:47-54); (disk_x, disk_y) = accepted random_in_unit_disk sample.
// direction = llc + s*hor + t*ver - origin - offset, with (llc - origin) folded on the host in f64.
template <typename T> __device__ __forceinline__ void get_ray(const CameraT<T>& cam, T s, T t, T disk_x, T disk_y, V3<T>* orig, V3<T>* dir)
{
    T rdx = cam.lens_radius * disk_x, rdy = cam.lens_radius * disk_y;
    V3<T> offset = cam.u * rdx + cam.v * rdy;
    *orig = cam.origin + offset;
    *dir = cam.llc_minus_origin + cam.horizontal * s + cam.vertical * t - offset;
}

// ------------------------------------------------------------------------------------------------
// Sphere::hit (sphere.rs:16-41), in the form the renderer uses.
//
// The renderer keeps ray directions normalised (dhat) and carries |dir| separately, so t here is in
// units of dhat; the reference's t is t_here / |dir| (Appendix C.3).  The discriminant is evaluated
// as r^2 - |oc - (oc.dhat/a) dhat|^2 (distance from the centre to the ray's line) instead of
// half_b^2 - a*c: same sign and same roots in exact arithmetic, but no cancellation between
// half_b^2 and a*|oc|^2 for distant spheres, which float needs to meet the 1e-5 parity bar.
// oc = centre - origin (the negative of sphere.rs:18, so tca = -half_b/a).
// Range test and root order are sphere.rs:28-34 verbatim: root == t_max is ACCEPTED.
// ------------------------------------------------------------------------------------------------
// sqrt for the f64 test of large-radius spheres inside the f32 renderer: MUFU.RSQ seed + one Newton
// step in f64 (relative error ~2e-14, far below what t needs) instead of the ~30-instruction IEEE sqrt.
__device__ __forceinline__ double sqrt_seeded(double x)
{
    x = fmax(x, 1e-30);
    double y = (double)rsqrtf((float)x);
    y = y * fma(-0.5 * x, y * y, 1.5);
    return x * y;
}
template <bool kSeeded> __device__ __forceinline__ float sqrt_sel(float x) { return kSeeded ? (x > 0.0f ? x * rsqrtf(x) : 0.0f) : sqrtf(x); }   // MUFU.RSQ: 2 ulp
template <bool kSeeded> __device__ __forceinline__ double sqrt_sel(double x) { return kSeeded ? sqrt_seeded(x) : sqrt(x); }

// inv_a = 1 / |dhat|^2 (dhat is unit only to rounding; hoisted per ray).
template <typename T, bool kSeededSqrt = false>
__device__ __forceinline__ bool sphere_roots(V3<T> oc, V3<T> dhat, T inv_a, T r2, T t_min, T t_max, T* root)
{
    T tca = dot(oc, dhat) * inv_a;
    V3<T> l = oc - dhat * tca;
    T disc = r2 - length_squared(l);
    if (disc < T(0)) return false;                   // sphere.rs:25
    T sq = sqrt_sel<kSeededSqrt>(disc * inv_a);
    T t = tca - sq;                                  // sphere.rs:28
    if (t < t_min || t_max < t) {
        t = tca + sq;                                // sphere.rs:30
        if (!(t >= t_min) || t_max < t) return false;   // sphere.rs:31-33; also drops NaN
    }
    *root = t;
    return true;
}

// HitRecord::new (shapes/mod.rs:20-30) for a sphere hit at p: outward = (p - c) / r  (sphere.rs:37,
// Div = multiply by the reciprocal), front_face = dir . outward < 0, normal flipped on back faces.
template <typename T> __device__ __forceinline__ void hit_record(V3<T> p, V3<T> center, T radius, V3<T> dir, V3<T>* normal, bool* front_face)
{
    V3<T> outward = (p - center) * (T(1) / radius);
    bool ff = dot(dir, outward) < T(0);
    *front_face = ff;
    *normal = ff ? outward : neg(outward);
}

// ------------------------------------------------------------------------------------------------
// Scatter::scatter x3 (materials.rs:22-30, 50-61, 77-104).  `sample` is the injected random
// vector: Lambertian — in-unit-sphere vector before normalisation; Metal — in-unit-sphere vector;
// Dialectric — sample.x = xi.  r_dir may have any length (the reference never normalises rays).
// ------------------------------------------------------------------------------------------------
enum { MAT_LAMBERTIAN = 0, MAT_METAL = 1, MAT_DIELECTRIC = 2 };

template <typename T> __device__ __forceinline__ T reflectance(T cosine, T ref_idx)                                                     // materials.rs:78-82
{
    T r0 = (T(1) - ref_idx) / (T(1) + ref_idx);
    r0 = r0 * r0;
    T m = T(1) - cosine;
    return r0 + (T(1) - r0) * (m * m * m * m * m);
}

// kUnit: the caller guarantees that r_dir is already a unit vector and that a Lambertian sample is one too (the
// renderer's rays and its inversion sampler), so the three unit_vector() calls of the reference are identities up to
// rounding and are skipped; the unit-level entry points pass kUnit = false and normalise exactly where the reference does.
template <typename T, bool kUnit = false>
__device__ __forceinline__ bool scatter(int kind, V3<T> albedo, T param, V3<T> r_dir, V3<T> normal, bool front_face, V3<T> sample,
                                        V3<T>* attenuation, V3<T>* out_dir)
{
    if (kind == MAT_LAMBERTIAN) {                                                                     // materials.rs:22-30
        V3<T> d = normal + (kUnit ? sample : unit_vector(sample));
        if (is_near_zero(d)) d = normal;
        *out_dir = d; *attenuation = albedo;
        return true;
    } else if (kind == MAT_METAL) {                                                                   // materials.rs:50-61
        V3<T> reflected = kUnit ? reflect(r_dir, normal) : unit_vector(reflect(r_dir, normal));
        V3<T> d = reflected + sample * param;          // drawn even when fuzz == 0 (materials.rs:53)
        *out_dir = d; *attenuation = albedo;
        return !(dot(d, normal) <= T(0));
    } else {                                                                                          // materials.rs:77-104
        T ratio = front_face ? T(1) / param : param;
        V3<T> ud = kUnit ? r_dir : unit_vector(r_dir);
        T cos_theta = min_t(T(1), -dot(ud, normal));
        T sin_theta = sqrt_t(T(1) - cos_theta * cos_theta);
        bool can_refract = ratio * sin_theta <= T(1);
        if (can_refract && reflectance(cos_theta, ratio) <= sample.x) *out_dir = refract(ud, normal, ratio);
        else *out_dir = reflect(ud, normal);
        *attenuation = mk<T>(T(1), T(1), T(1));
        return true;
    }
}

// miss branch of ray_color (main.rs:54-56)
template <typename T, bool kUnit = false> __device__ __forceinline__ V3<T> sky(V3<T> dir)
{
    V3<T> ud = kUnit ? dir : unit_vector(dir);
    T t = T(0.5) * (ud.y + T(1));
    return mk<T>(T(1), T(1), T(1)) * (T(1) - t) + mk<T>(T(0.5), T(0.7), T(1.0)) * t;
}

// ------------------------------------------------------------------------------------------------
// packed f32x2 arithmetic (sm_100 FFMA2 / FADD2 / FMUL2): one issue slot, two FP32 lanes-worth of
// work.  The scan's filter runs on these so that the loads, sign-extraction and loop control fit
// in the issue slots the packed ops leave free (tools/probe_fp32.cu, profiles/).
// ------------------------------------------------------------------------------------------------
typedef unsigned long long u64;
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) { float2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b), "l"(*(u64*)&c)); return d; }
__device__ __forceinline__ float2 fadd2(float2 a, float2 b) { float2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 fsub2(float2 a, float2 b) { float2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 fmul2(float2 a, float2 b) { float2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(*(u64*)&d) : "l"(*(u64*)&a), "l"(*(u64*)&b)); return d; }
__device__ __forceinline__ float2 neg2(float2 a) { return make_float2(-a.x, -a.y); }
__device__ __forceinline__ float2 bc2(float a) { return make_float2(a, a); }

}  // namespace rt
