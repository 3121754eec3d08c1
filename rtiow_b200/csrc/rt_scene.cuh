// rt_scene.cuh — device-resident scene layout and the closest-hit scan (HittableList::hit,
// /root/reference/src/shapes/mod.rs:56-69) of the B200 render path.
//
// Layout in HBM (built once by rtiow_scene_upload, read-only afterwards):
//   soa      float [4][np]   cx | cy | cz | K     "small" spheres in list order, padded to a multiple
//                            of 32 with never-hit entries; staged into shared memory by each CTA and
//                            streamed as broadcast LDS.128 (4 spheres per load per array)
//   small    float4 [np]     (cx, cy, cz, r) exact f32 copy for the precise test of candidates
//   small_idx int   [np]     position in `small` -> index in the reference's list (-1 = padding)
//   big      double4 [nb]    (cx, cy, cz, r) spheres whose radius makes |oc|^2 - r^2 cancel in f32
//                            (the r=1000 ground, main.rs:64); tested in f64, 1 of ~530 tests
//   big_idx  int    [nb]
//   sph      float4 / double4 [n]  (cx,cy,cz,r) in list order, for HitRecord::new
//   mat      float4 / double4 [n]  (albedo r,g,b, fuzz|ir) resolved per sphere; kind uint8 [n]
#pragma once
#include "rt_device.cuh"

namespace rt {

struct SceneDev {
    const float* soa;
    const float4* small;
    const int* small_idx;
    int np;
    int w_cull;             // 32-sphere words [0, w_cull) may use behind-the-ray culling (rt_scene.cuh, filter_word)
    const double4* big;
    const int* big_idx;
    int nb;
    const float4* sph;
    const double4* sphd;
    const float4* mat;
    const double4* matd;
    const uint8_t* kind;
    int n;
};

#define RT_FULL 0xffffffffu
// Conservative slack of the f32 filter, in units of (|c|^2 + |o|^2): 96 * 2^-24.  The filter's rounding
// error is bounded by ~40 u (|c|^2 + |o|^2) + 4 u r^2 (u = 2^-24; derivation in DESIGN.md), so with this
// slack every sphere the precise test can accept passes the filter.  The per-sphere share is folded
// into K at upload, the per-ray share into |o|^2 below: no cost inside the loop.
#define RT_FILTER_SLACK 5.7220458984375e-06f
#define RT_CAND_CAP 16           // candidate slots per lane per segment (uint16 positions)
#define RT_SEG_WORDS 32          // 32 words x 32 spheres per segment between drains

struct HitF {           // closest hit so far; t in units of the normalised direction
    float t;
    int idx;            // index in the reference's list, -1 = none
    int code;           // where the winner lives: >= 0 position in `small`, <= -2 : -2 - position in `big`, -1 none
};
#define RT_SELF_NONE (-1)

// Precise test of one candidate (sphere.rs:16-34 via sphere_roots), index-aware acceptance.
// The reference scans in list order and accepts root <= closest_so_far, so among equal roots the
// LARGEST index wins (sphere.rs:29,31; mod.rs:61-66); `t < best || (t == best && idx > best_idx)`
// gives the same winner for any processing order.
template <typename T, bool kSeededSqrt = false>
__device__ __forceinline__ void candidate(V3<T> o, V3<T> dhat, T inv_a, T t_min, V3<T> c, T r, int idx, T* t_best, int* i_best)
{
    T t;
    // t_max = +inf: a root beyond the current closest is rejected by the comparison below, exactly
    // as sphere.rs:29-33 rejects it (its far root is farther still)
    if (!sphere_roots<T, kSeededSqrt>(c - o, dhat, inv_a, r * r, t_min, (T)__int_as_float(0x7f800000), &t)) return;
    if (t < *t_best || (t == *t_best && idx > *i_best)) { *t_best = t; *i_best = idx; }
}

// The sphere a ray STARTS on (the one it just scattered from).  In f64 the reference's origin is within
// ~1e-15 of that surface, so its near-zero root never reaches t_min = 1e-4 (main.rs:44); an f32 origin
// is ~1e-6 off, which t_min * |dir| does not cover when the scattered direction is short
// (Lambertian normal + unit vector, materials.rs:23).  For that one sphere the origin is therefore
// taken ON the surface (centre + self_n * |r|, self_n = unit(p - centre)): the roots are exactly
// {0, 2 tca}, the zero root is dropped as sphere.rs:29 drops it, and only 2 tca is offered.
template <typename T>
__device__ __forceinline__ void candidate_self(V3<T> dhat, T inv_a, T t_min, V3<T> self_n, T r, int idx, T* t_best, int* i_best)
{
    const T t = T(-2) * abs_t(r) * dot(self_n, dhat) * inv_a;      // oc = -|r| self_n ; t = 2 (oc . dhat) / a
    if (!(t >= t_min)) return;
    if (t < *t_best || (t == *t_best && idx > *i_best)) { *t_best = t; *i_best = idx; }
}

// The f32 scan over the shared-memory SoA.
//   filter (all spheres, branch-free, packed f32x2 over sphere pairs), sphere.rs:18-25 expanded around
//   the coordinate origin so that the per-ray and per-sphere parts separate:
//       hb = c.d - o.d ;  C = K + |o|^2 - 2 c.o ,  K = |c|^2 - r^2 ;  disc' = hb^2 - C     (8 packed ops per 2 tests)
//     With K and |o|^2 lowered by RT_FILTER_SLACK, disc' >= 0 is a conservative superset of
//     "discriminant >= 0" (sphere.rs:24-25); its sign bit is funnel-shifted into a 32-sphere word
//     (1 SHF per test on the ALU pipe, co-issued).  The expansion cancels |c|^2 + |o|^2 against 2 c.o, which
//     is why it is only a filter: every survivor goes through the well-conditioned sphere_roots.
//   candidates (a few per ray): positions appended to a per-lane list in shared memory, then the
//     whole warp drains its lists in lock-step through candidate<float>.
// s_soa: [4][np] floats in shared memory (or global when the scene does not fit).
// Per-ray constants of the filter (rt_scene.cuh header comment): broadcast scalars for the packed ops.
struct FilterRay {
    float2 M2OX, M2OY, M2OZ, DX, DY, DZ, NOD, OO;
};
__device__ __forceinline__ FilterRay make_filter_ray(V3<float> o, V3<float> dhat)
{
    FilterRay f;
    const float oo = length_squared(o);
    f.M2OX = bc2(-2.0f * o.x); f.M2OY = bc2(-2.0f * o.y); f.M2OZ = bc2(-2.0f * o.z);
    f.DX = bc2(dhat.x); f.DY = bc2(dhat.y); f.DZ = bc2(dhat.z);
    f.NOD = bc2(-dot(o, dhat));
    f.OO = bc2(oo - RT_FILTER_SLACK * oo);
    return f;
}

// One 32-sphere word of the filter.  kCull: the last op is hb*|hb| - C instead of hb*hb - C (the |.| is an operand
// modifier of FFMA2, so it is free): a sphere whose centre lies BEHIND the ray (hb < 0) then only passes if the origin
// is deep inside it, i.e. spheres entirely behind an outside origin — about half of all line/sphere intersections of a
// bounce ray — never become candidates.  Valid only for spheres no ray origin can be inside of (see `n_cull` below).
template <bool kCull>
__device__ __forceinline__ unsigned filter_word(const float4* __restrict__ CX, const float4* __restrict__ CY, const float4* __restrict__ CZ,
                                                const float4* __restrict__ KK, int q0, int q_next_word, const FilterRay& f,
                                                float4& ncx, float4& ncy, float4& ncz, float4& nkk)
{
    unsigned m = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        const float4 cx = ncx, cy = ncy, cz = ncz, kk = nkk;
        const int qn = q < 7 ? q0 + q + 1 : q_next_word;
        ncx = CX[qn]; ncy = CY[qn]; ncz = CZ[qn]; nkk = KK[qn];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const float2 X = h ? make_float2(cx.z, cx.w) : make_float2(cx.x, cx.y);
            const float2 Y = h ? make_float2(cy.z, cy.w) : make_float2(cy.x, cy.y);
            const float2 Z = h ? make_float2(cz.z, cz.w) : make_float2(cz.x, cz.y);
            const float2 K = h ? make_float2(kk.z, kk.w) : make_float2(kk.x, kk.y);
            const float2 hb = ffma2(X, f.DX, ffma2(Y, f.DY, ffma2(Z, f.DZ, f.NOD)));
            const float2 C = ffma2(X, f.M2OX, ffma2(Y, f.M2OY, ffma2(Z, f.M2OZ, fadd2(K, f.OO))));
            const float2 hb2 = kCull ? make_float2(fabsf(hb.x), fabsf(hb.y)) : hb;
            const float2 disc = ffma2(hb, hb2, neg2(C));
            m = __funnelshift_l(__float_as_uint(disc.x), m, 1);
            m = __funnelshift_l(__float_as_uint(disc.y), m, 1);
        }
    }
    return m;
}

// words [0, w_cull) hold spheres that provably contain no ray origin (they overlap no other sphere and not the
// camera lens): filtered with behind-the-ray culling; words [w_cull, n_words) hold the rest, filtered without.
template <bool kSmem>
__device__ __forceinline__ void scan_small(const float* __restrict__ s_soa, int np, int w_cull, const float4* __restrict__ small,
                                           V3<float> o, V3<float> dhat, float inv_a, float t_min, int self_pos, V3<float> self_n,
                                           uint16_t* cand, int cand_stride, float* t_best, int* p_best)
{
    const int n4 = np >> 2;
    const float4* CX = reinterpret_cast<const float4*>(s_soa);
    const float4* CY = CX + n4;
    const float4* CZ = CY + n4;
    const float4* KK = CZ + n4;
    const FilterRay f = make_filter_ray(o, dhat);
    const int n_words = np >> 5;
    int pb = *p_best; float tb = *t_best;
    // the sphere the ray starts on is tested on its own, independently of the filter
    if (self_pos >= 0) candidate_self<float>(dhat, inv_a, t_min, self_n, small[self_pos].w, self_pos, &tb, &pb);

    // software pipeline: the quad (4 spheres x 4 arrays, 4 LDS.128) for step q+1 is loaded while step q
    // computes, so the ~30-cycle shared-memory latency hides behind 16 FFMA2 of the same warp
    if (n_words == 0) { *p_best = pb; *t_best = tb; return; }
    float4 ncx = CX[0], ncy = CY[0], ncz = CZ[0], nkk = KK[0];
    for (int w0 = 0; w0 < n_words; w0 += RT_SEG_WORDS) {
        const int w1 = min(w0 + RT_SEG_WORDS, n_words);
        int nc = 0;
        for (int w = w0; w < w1; ++w) {
            const int q0 = w << 3;
            const int q_next_word = (w + 1 < n_words) ? q0 + 8 : q0;     // the last word re-reads its first quad (unused)
            const unsigned m = w < w_cull ? filter_word<true>(CX, CY, CZ, KK, q0, q_next_word, f, ncx, ncy, ncz, nkk)
                                          : filter_word<false>(CX, CY, CZ, KK, q0, q_next_word, f, ncx, ncy, ncz, nkk);
            unsigned c = ~m;                         // bit (31-k) set: sphere 32w+k passed the filter
            while (c) {
                const int k = __clz(c);
                c &= ~(0x80000000u >> k);
                const int p = (w << 5) + k;
                if (nc < RT_CAND_CAP) { cand[nc * cand_stride] = (uint16_t)(p - (w0 << 5)); ++nc; }
                else if (p != self_pos) { const float4 s = small[p]; candidate<float, true>(o, dhat, inv_a, t_min, mk(s.x, s.y, s.z), s.w, p, &tb, &pb); }
            }
        }
        const int nmax = __reduce_max_sync(RT_FULL, nc);
        for (int k = 0; k < nmax; ++k) {
            if (k < nc) {
                const int p = (w0 << 5) + cand[k * cand_stride];
                if (p != self_pos) { const float4 s = small[p]; candidate<float, true>(o, dhat, inv_a, t_min, mk(s.x, s.y, s.z), s.w, p, &tb, &pb); }
            }
        }
    }
    *p_best = pb; *t_best = tb;
}

// Closest hit of one ray against the whole scene: HittableList::hit (mod.rs:56-69).
// float: packed filter over the small spheres + f64 test of the big ones; all lanes of the warp
// must call together.  self_code / self_n identify the sphere the ray starts on (RT_SELF_NONE for
// camera rays).  Returns t in dhat units, the list index (or -1) and the winner's code.
template <bool kSmem>
__device__ __forceinline__ HitF closest_hit(const SceneDev& sc, const float* s_soa, V3<float> o, V3<float> dhat, float t_min,
                                            int self_code, V3<float> self_n, uint16_t* cand, int cand_stride)
{
    const float inv_a = 1.0f / length_squared(dhat);
    float tb = __int_as_float(0x7f800000);   // f64::INFINITY at main.rs:44
    int pb = -1;
    scan_small<kSmem>(s_soa, sc.np, sc.w_cull, sc.small, o, dhat, inv_a, t_min, self_code, self_n, cand, cand_stride, &tb, &pb);
    HitF h; h.t = tb; h.idx = pb >= 0 ? sc.small_idx[pb] : -1; h.code = pb;
    if (sc.nb > 0) {
        const V3<double> od = mk<double>(o.x, o.y, o.z), dd = mk<double>(dhat.x, dhat.y, dhat.z);
        const double inv_ad = 2.0 - length_squared(dd);          // 1/a for a = 1 + e, |e| < 1e-6: exact to e^2
        double tbd = (double)h.t; int ib = h.idx;
        for (int b = 0; b < sc.nb; ++b) {
            const double4 s = sc.big[b];
            const int before = ib; const double tbefore = tbd;
            if (self_code == -2 - b) {
                V3<double> sn = mk<double>(self_n.x, self_n.y, self_n.z);
                sn = sn * (1.0 / sqrt(length_squared(sn)));
                candidate_self<double>(dd, inv_ad, (double)t_min, sn, s.w, sc.big_idx[b], &tbd, &ib);
            } else {
                candidate<double, true>(od, dd, inv_ad, (double)t_min, mk(s.x, s.y, s.z), s.w, sc.big_idx[b], &tbd, &ib);
            }
            if (ib != before || tbd != tbefore) { h.idx = ib; h.t = (float)tbd; h.code = -2 - b; }
        }
    }
    return h;
}

// double: the reference's arithmetic width, every sphere through the precise test (triage mode).
// code == list index here.
__device__ __forceinline__ void closest_hit_f64(const SceneDev& sc, V3<double> o, V3<double> dhat, double t_min, int self_idx, V3<double> self_n,
                                                double* t_out, int* idx_out)
{
    const double inv_a = 1.0 / length_squared(dhat);
    double tb = __longlong_as_double(0x7ff0000000000000LL);
    int ib = -1;
    for (int i = 0; i < sc.n; ++i) {
        const double4 s = sc.sphd[i];
        if (i == self_idx) candidate_self<double>(dhat, inv_a, t_min, self_n, s.w, i, &tb, &ib);
        else candidate<double>(o, dhat, inv_a, t_min, mk(s.x, s.y, s.z), s.w, i, &tb, &ib);
    }
    *t_out = tb; *idx_out = ib;
}

// cooperative copy of the filter SoA into shared memory (16-byte vector copies; np % 32 == 0)
__device__ __forceinline__ void stage_scene(float* s_soa, const float* __restrict__ g_soa, int np)
{
    const float4* src = reinterpret_cast<const float4*>(g_soa);
    float4* dst = reinterpret_cast<float4*>(s_soa);
    for (int i = threadIdx.x; i < np; i += blockDim.x) dst[i] = src[i];   // 4*np floats = np float4
    __syncthreads();
}

}  // namespace rt
