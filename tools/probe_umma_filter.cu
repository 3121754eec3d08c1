// tools/probe_umma_filter.cu — stand-alone probe of the sphere filter as a tensor-core contraction.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -I rtiow_b200/csrc -o tools/probe_umma_filter tools/probe_umma_filter.cu
//
// The filter of rt_scene.cuh (7 FFMA2 per sphere pair on the FP32 pipe, 65.8 TFLOP/s-equivalent stand-alone) restated as
// D[128 rays x N spheres] = A[128 x 11 features] . B[N x 11]^T with a 3-product fp16 hi/lo split (rt_umma.cuh), A written to
// TMEM by the threads, B static in shared memory, D read back with tcgen05.ld and the sign bits funnel-shifted into 32-sphere
// words.  Reports, one JSON line each:
//   * accuracy against the f64 discriminant of sphere.rs:18-25 (max |D - disc|, false negatives at the chosen slack,
//     false-candidate rate against the exact test), for both readings of the descriptor's LBO/SBO fields;
//   * tests/s and its 17-FLOP-per-test equivalent for several (groups per CTA, spheres per chunk) shapes, with and without
//     the MMA and with and without the sign collection, so the MMA floor, the TMEM read floor and the ALU floor show separately.
// Kill criterion (VERDICT): ship in situ only if the best full configuration beats 65.8 TFLOP/s-equivalent by >= 1.5x.
// Not part of the product path.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <vector>
#include <algorithm>
#include <cuda_runtime.h>
#include "rt_umma.cuh"

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { fprintf(stderr, "CUDA %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1); } } while (0)

using namespace rt::umma;

struct ProbeArgs {
    const unsigned char* bimg;   // [2][npad * 32] bytes: B_hi block, B_lo block (canonical layout)
    const float4* rays_o;        // per ray: origin (w unused)
    const float4* rays_d;        // per ray: unit direction
    int npad;                    // spheres incl. padding: a multiple of the chunk size
    int iters;
    float R2;                    // squared bounding radius of the table's spheres
    FeatScale sc;
    float* dump;                 // [dump_rays][npad]
    int dump_rays;
    unsigned long long* out;     // [0] candidates (sign clear), [1] xor of the masks
    long long* trace;            // F_TRACE: clock64 stamps of block 0, group 0, warp 0, lane 0 during scan `iters - 1`
};

enum { F_MMA = 1, F_SIGN = 2, F_DUMP = 4, F_SWAP = 8, F_NOLD = 16, F_TRACE = 32, F_SIGN3 = 64, F_D16 = 128 };

// F_D16: the accumulator in fp16 (idesc c_format = F16).  Only the SIGN of D is used, and rounding the final sum to fp16 keeps
// it; the cross terms go first (they are small: fp16 rounding of the intermediate D costs ~1e-4) and hi.hi last.  D then comes
// back as packed halves (tcgen05.ld ... .pack::16b: 32 spheres in 16 registers) and the sign bits are collected FOUR per
// instruction: PRMT with sign replication turns two registers into four sign bytes, a LOP3 bit-select tree interleaves eight
// such words into one 32-sphere mask: 8 PRMT + 7 LOP3 per word instead of 32 SHF.
// make_idesc_f16_f16, tmem_ld16p, sign_word16: rt_umma.cuh (the product kernels use them since this probe passed)

template <int G, int NC, int NBUF, int FLAGS>
__global__ void __launch_bounds__(G * 160, 1) probe_kernel(const ProbeArgs a)
{
    static_assert(NC % 32 == 0 && NC >= 32 && NC <= 256, "chunk = whole 32-sphere words");
    static_assert(NBUF == 1 || NBUF == 2, "D buffers per group");
    constexpr int W = NBUF * NC + 16;                // TMEM columns per group: D buffers + A_hi (8) + A_lo (8)
    static_assert(G * W <= 512, "TMEM has 512 columns");
    extern __shared__ __align__(1024) unsigned char smem[];
    // warps [0, 4G): ray warps, group g = warp / 4, TMEM lane quarter q = warp % 4; warps [4G, 5G): the groups' MMA issuers
    // (tcgen05.mma issue blocks the issuing thread while the tensor pipe is busy, so it cannot share a warp with an epilogue)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool issuer_warp = warp >= 4 * G;
    const int g = issuer_warp ? warp - 4 * G : warp >> 2, q = warp & 3;
    const size_t blk = RT_UMMA_B_BLOCK_BYTES(a.npad);
    unsigned char* b_hi = smem;
    unsigned char* b_lo = smem + blk;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * blk);            // [G][8]: a_full, full0, full1, empty0, empty1
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G * 8);

    for (size_t i = (size_t)tid * 16; i < 2 * blk; i += (size_t)blockDim.x * 16)
        *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(a.bimg + i);
    if (tid == 0) {
        for (int i = 0; i < G; ++i) {
            mbar_init(smem_u32(bars + 8 * i + 0), 128);
            mbar_init(smem_u32(bars + 8 * i + 1), 1); mbar_init(smem_u32(bars + 8 * i + 2), 1);
            mbar_init(smem_u32(bars + 8 * i + 3), 4); mbar_init(smem_u32(bars + 8 * i + 4), 4);
        }
        fence_mbar_init();
    }
    __syncwarp();
    if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    fence_proxy_async_smem();                        // B was written with generic stores, the tensor core reads it through the async proxy
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    const uint32_t col0 = tmem_base + (uint32_t)(g * W);
    const uint32_t t_ahi = col0 + NBUF * NC, t_alo = t_ahi + 8;
    const uint32_t bar_afull = smem_u32(bars + 8 * g), bar_full0 = smem_u32(bars + 8 * g + 1), bar_empty0 = smem_u32(bars + 8 * g + 3);   // buffer b: + 8 b
    const int n_chunks = a.npad / NC;
    unsigned long long cand = 0; unsigned xr = 0;

    if (issuer_warp) {
        if (lane == 0 && (FLAGS & F_MMA)) {
            const uint32_t idesc = (FLAGS & F_D16) ? make_idesc_f16_f16(NC) : make_idesc_f16_f32(NC);
            const uint32_t lbo = (FLAGS & F_SWAP) ? RT_UMMA_B_SBO : RT_UMMA_B_LBO, sbo = (FLAGS & F_SWAP) ? RT_UMMA_B_LBO : RT_UMMA_B_SBO;
            const uint32_t s_hi = smem_u32(b_hi), s_lo = smem_u32(b_lo);
            uint32_t a_phase = 0, use_phase = 0, used = 0;
            for (int it = 0; it < a.iters; ++it) {
                mbar_wait(bar_afull, a_phase); a_phase ^= 1u;                 // all 128 rays' feature rows are in TMEM
                tc_fence_after();
                for (int c = 0; c < n_chunks; ++c) {
                    const uint32_t b = NBUF == 1 ? 0u : (uint32_t)(c & 1);
                    if ((used >> b) & 1u) { mbar_wait(bar_empty0 + 8u * b, (use_phase >> b) & 1u); use_phase ^= 1u << b; tc_fence_after(); }
                    used |= 1u << b;
                    const uint32_t t_d = col0 + b * NC;
                    const uint32_t off = (uint32_t)(c * (NC / 8)) * 256u;
                    const uint64_t dh = make_smem_desc(s_hi + off, lbo, sbo), dl = make_smem_desc(s_lo + off, lbo, sbo);
                    if (FLAGS & F_D16) {
                        mma_f16_ts(t_d, t_ahi, dl, idesc, 0u);               // hi . lo
                        mma_f16_ts(t_d, t_alo, dh, idesc, 1u);               // lo . hi
                        mma_f16_ts(t_d, t_ahi, dh, idesc, 1u);               // hi . hi last: the final rounding to fp16 keeps the sign
                    } else {
                        mma_f16_ts(t_d, t_ahi, dh, idesc, 0u);
                        mma_f16_ts(t_d, t_ahi, dl, idesc, 1u);
                        mma_f16_ts(t_d, t_alo, dh, idesc, 1u);
                    }
                    tc_commit(bar_full0 + 8u * b);
                }
            }
        }
    } else {
        const uint32_t lane_base = (uint32_t)(32 * q) << 16;
        const int ray = (blockIdx.x * G + g) * 128 + q * 32 + lane;
        const float4 ro = a.rays_o[ray], rd = a.rays_d[ray];
        float ox = ro.x, oy = ro.y, oz = ro.z;
        const float dx = rd.x, dy = rd.y, dz = rd.z;
        uint32_t full_phase = 0;                     // bit b: parity of the next phase of full[b] to wait for
        const bool tracer = (FLAGS & F_TRACE) && blockIdx.x == 0 && tid == 0;
        int tn = 0;
#define STAMP() do { if (tracer && it == a.iters - 1 && tn < 64) a.trace[tn++] = clock64(); } while (0)
        for (int it = 0; it < a.iters; ++it) {
            STAMP();
            // the ray's feature rows -> TMEM (one row = one lane = one ray)
            const float ia = 2.0f - (dx * dx + dy * dy + dz * dz);
            float t = (ox * dx + oy * dy + oz * dz) * ia;
            float fx = ox - t * dx, fy = oy - t * dy, fz = oz - t * dz;
            t = (fx * dx + fy * dy + fz * dz) * ia;
            fx -= t * dx; fy -= t * dy; fz -= t * dz;
            const bool live = fx * fx + fy * fy + fz * fz < a.R2;
            uint32_t hi[8], lo[8];
            ray_features(fx, fy, fz, dx, dy, dz, live, 0.0f, a.sc, hi, lo);
            tmem_st8(t_ahi + lane_base, hi);
            tmem_st8(t_alo + lane_base, lo);
            tc_wait_st();
            tc_fence_before();
            if (FLAGS & F_MMA) mbar_arrive(bar_afull);
            STAMP();
            for (int c = 0; c < n_chunks; ++c) {
                const uint32_t b = NBUF == 1 ? 0u : (uint32_t)(c & 1);
                if (FLAGS & F_MMA) { mbar_wait(bar_full0 + 8u * b, (full_phase >> b) & 1u); full_phase ^= 1u << b; }
                STAMP();
                tc_fence_after();
                constexpr int NW = NC / 32;
                if (FLAGS & F_D16) {
                    uint32_t v[NW][16];
#pragma unroll
                    for (int w = 0; w < NW; ++w) tmem_ld16p(col0 + b * NC + lane_base + 32 * w, v[w]);
                    tc_wait_ld();
                    STAMP();
                    if (FLAGS & F_MMA) { tc_fence_before(); __syncwarp(); if (lane == 0) mbar_arrive(bar_empty0 + 8u * b); }
                    if ((FLAGS & F_DUMP) && it == 0 && ray < a.dump_rays) {
#pragma unroll
                        for (int w = 0; w < NW; ++w)
#pragma unroll
                            for (int k = 0; k < 16; ++k) {
                                const __half2 h = *reinterpret_cast<const __half2*>(&v[w][k]);
                                a.dump[(size_t)ray * a.npad + c * NC + w * 32 + 2 * k] = __low2float(h);
                                a.dump[(size_t)ray * a.npad + c * NC + w * 32 + 2 * k + 1] = __high2float(h);
                            }
                    }
#pragma unroll
                    for (int w = 0; w < NW; ++w) {
                        if (FLAGS & F_SIGN) {
                            const unsigned m = sign_word16(v[w]); cand += __popc(~m);
                            if (FLAGS & F_DUMP) {                             // self-check of the bit order: bit 8 b + j = sign of half 4 j + b
                                unsigned ref = 0;
#pragma unroll
                                for (int h = 0; h < 32; ++h) ref |= ((v[w][h >> 1] >> ((h & 1) ? 31 : 15)) & 1u) << (8 * (h & 3) + (h >> 2));
                                xr += (m != ref);
                            } else xr ^= m;
                        }
                        else {
#pragma unroll
                            for (int k = 0; k < 16; ++k) asm volatile("" ::"r"(v[w][k]));
                        }
                    }
                    continue;
                }
                // SIGN3 needs its three words together; otherwise two words (64 registers) are in flight at a time
                constexpr int SUB = (FLAGS & F_SIGN3) ? 3 : (NW >= 2 ? 2 : 1);
#pragma unroll
                for (int w0 = 0; w0 < NW; w0 += SUB) {
                    uint32_t v[SUB][32];
#pragma unroll
                    for (int w = 0; w < SUB; ++w) {
                        if (w0 + w >= NW) break;
                        if (!(FLAGS & F_NOLD)) tmem_ld32(col0 + b * NC + lane_base + 32 * (w0 + w), v[w]);
                        else {
#pragma unroll
                            for (int k = 0; k < 32; ++k) v[w][k] = 0x80000000u;
                        }
                    }
                    if (!(FLAGS & F_NOLD)) tc_wait_ld();
                    if (w0 + SUB >= NW) {                                  // the chunk's last loads are done: hand the D buffer back
                        STAMP();
                        if (FLAGS & F_MMA) {
                            tc_fence_before();
                            __syncwarp();
                            if (lane == 0) mbar_arrive(bar_empty0 + 8u * b);
                        }
                    }
                    if ((FLAGS & F_DUMP) && it == 0 && ray < a.dump_rays) {
#pragma unroll
                        for (int w = 0; w < SUB; ++w)
#pragma unroll
                            for (int k = 0; k < 32; ++k)
                                if (w0 + w < NW) a.dump[(size_t)ray * a.npad + c * NC + (w0 + w) * 32 + k] = __uint_as_float(v[w][k]);
                    }
                    if (FLAGS & F_SIGN3) {
                        // three words at a time: AND the three sign bits of position k (one LOP3), then one SHF: bit (31-k) is CLEAR when
                        // any of the three spheres {k, 32+k, 64+k} passed; the (rare) clear bits are resolved afterwards
                        static_assert(!(FLAGS & F_SIGN3) || NW % 3 == 0, "SIGN3 needs whole triples of words");
                        unsigned m = 0;
#pragma unroll
                        for (int k = 0; k < 32; ++k) m = __funnelshift_l(v[0][k] & v[1 % SUB][k] & v[2 % SUB][k], m, 1);
                        unsigned clear = ~m;
                        while (clear) {                                   // rare: ~1 bit per ray per scan
                            const int k = __clz(clear); clear &= ~(0x80000000u >> k);
                            unsigned which = 0;
#pragma unroll
                            for (int kk = 0; kk < 32; ++kk)
                                if (kk == k) which = ((v[0][kk] >> 31) ^ 1u) | (((v[1 % SUB][kk] >> 31) ^ 1u) << 1) | (((v[2 % SUB][kk] >> 31) ^ 1u) << 2);
                            cand += __popc(which); xr ^= which << k;
                        }
                    } else {
#pragma unroll
                        for (int w = 0; w < SUB; ++w) {
                            if (w0 + w >= NW) break;
                            if (FLAGS & F_SIGN) {
                                unsigned m = 0;
#pragma unroll
                                for (int k = 0; k < 32; ++k) m = __funnelshift_l(v[w][k], m, 1);      // bit (31-k) = sign of sphere k's discriminant
                                cand += __popc(~m); xr ^= m;
                            } else {
#pragma unroll
                                for (int k = 0; k < 32; ++k) asm volatile("" ::"r"(v[w][k]));         // the loads stay, nothing else
                            }
                        }
                    }
                }
            }
            ox += 1e-3f * dx; oy += 1e-3f * dy; oz += 1e-3f * dz;      // same line, new numbers
        }
        atomicAdd(a.out, cand);
        atomicXor(a.out + 1, (unsigned long long)xr);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// ---- raw tcgen05.mma issue/execute rate: one thread issues `reps` MMAs back to back, no epilogue ---------------------------
// SAME_D: all into one accumulator (a K loop) / alternating between two.  Reports cycles per MMA to issue and to complete.
template <int NC, int SAME_D, int TS>
__global__ void __launch_bounds__(128, 1) mma_rate_kernel(const unsigned char* bimg, int npad, int reps, long long* out)
{
    extern __shared__ __align__(1024) unsigned char smem[];
    const int tid = threadIdx.x, warp = tid >> 5;
    const size_t blk = RT_UMMA_B_BLOCK_BYTES(npad);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 2 * blk);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2);
    unsigned char* a_smem = smem + 2 * blk + 1024;           // 128 rows x 16 fp16, canonical K-major no-swizzle: 4 KB
    for (size_t i = (size_t)tid * 16; i < 2 * blk; i += 128 * 16) *reinterpret_cast<uint4*>(smem + i) = *reinterpret_cast<const uint4*>(bimg + i);
    for (int i = tid; i < 1024; i += 128) reinterpret_cast<uint32_t*>(a_smem)[i] = 0x3c003c00u;   // fp16 1.0
    if (tid == 0) { mbar_init(smem_u32(bars), 1); fence_mbar_init(); }
    __syncwarp();
    if (warp == 0) { tmem_alloc(smem_u32(tmem_slot), 512); tmem_relinquish(); }
    fence_proxy_async_smem();
    tc_fence_before(); __syncthreads(); tc_fence_after();
    const uint32_t tb = *tmem_slot;
    constexpr uint32_t kACol = SAME_D ? NC : 2 * NC;
    static_assert(kACol + 8 <= 512, "TMEM columns");
    uint32_t ones[8]; for (int i = 0; i < 8; ++i) ones[i] = 0x3c003c00u;
    tmem_st8(tb + kACol + ((uint32_t)(32 * warp) << 16), ones);
    tc_wait_st(); tc_fence_before(); __syncthreads();
    if (tid == 0) {
        tc_fence_after();
        const uint32_t idesc = make_idesc_f16_f32(NC);
        const uint64_t db = make_smem_desc(smem_u32(smem), RT_UMMA_B_LBO, RT_UMMA_B_SBO);
        const uint64_t da = make_smem_desc(smem_u32(a_smem), 128u, 256u);
        const long long t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t d = tb + ((SAME_D || !(r & 1)) ? 0u : (uint32_t)NC);
            if (TS) mma_f16_ts(d, tb + kACol, db, idesc, r > 1 ? 1u : 0u);
            else {
                asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                             ::"r"(d), "l"(da), "l"(db), "r"(idesc), "r"(r > 1 ? 1u : 0u) : "memory");
            }
        }
        tc_commit(smem_u32(bars));
        const long long t1 = clock64();
        mbar_wait(smem_u32(bars), 0);
        const long long t2 = clock64();
        out[0] = t1 - t0; out[1] = t2 - t0;
    }
    tc_fence_before(); __syncthreads();
    if (warp == 0) tmem_dealloc(tb, 512);
}

template <int NC, int SAME_D, int TS>
static void run_rate(const unsigned char* d_b, int npad, long long* d_out, size_t smem_bytes)
{
    auto k = mma_rate_kernel<NC, SAME_D, TS>;
    CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    for (int reps : {64, 1024}) {
        k<<<1, 128, smem_bytes>>>(d_b, npad, reps, d_out);
        CK(cudaDeviceSynchronize());
        long long o[2]; CK(cudaMemcpy(o, d_out, 16, cudaMemcpyDeviceToHost));
        printf("{\"probe\":\"umma_rate\",\"n\":%d,\"same_d\":%d,\"a_in_tmem\":%d,\"reps\":%d,\"issue_cycles_per_mma\":%.1f,\"complete_cycles_per_mma\":%.1f,\"floor\":%d}\n",
               NC, SAME_D, TS, reps, (double)o[0] / reps, (double)o[1] / reps, NC / 2);
    }
}

// ---- host --------------------------------------------------------------------------------------------------------------
static unsigned long long g_rng = 0x9E3779B97F4A7C15ull;
static double urand() { g_rng = g_rng * 6364136223846793005ull + 1442695040888963407ull; return (double)(g_rng >> 11) * (1.0 / 9007199254740992.0); }

struct Sph { double x, y, z, r; };

static unsigned short f2h(float f) { __half h = __float2half_rn(f); unsigned short u; memcpy(&u, &h, 2); return u; }
static float h2f(unsigned short u) { __half h; memcpy(&h, &u, 2); return __half2float(h); }

template <int G, int NC, int NBUF, int FLAGS>
static void run(const char* name, ProbeArgs a, int n_real, size_t smem_bytes, int sms)
{
    const int flags = FLAGS;
    CK(cudaFuncSetAttribute(probe_kernel<G, NC, NBUF, FLAGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
    CK(cudaMemset(a.out, 0, 16));
    ProbeArgs warm = a; warm.iters = 2;
    probe_kernel<G, NC, NBUF, FLAGS><<<sms, G * 160, smem_bytes>>>(warm);
    CK(cudaDeviceSynchronize());
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        CK(cudaMemset(a.out, 0, 16));
        CK(cudaEventRecord(e0));
        probe_kernel<G, NC, NBUF, FLAGS><<<sms, G * 160, smem_bytes>>>(a);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); best = std::min(best, ms);
    }
    unsigned long long out[2]; CK(cudaMemcpy(out, a.out, 16, cudaMemcpyDeviceToHost));
    const double rays = (double)sms * G * 128 * a.iters;
    const double tests = rays * n_real;                          // padding spheres are not counted as work
    printf("{\"probe\":\"umma_filter\",\"name\":\"%s\",\"groups\":%d,\"chunk\":%d,\"d_buffers\":%d,\"mma\":%d,\"sign\":%d,\"ms\":%.4f,\"Ttests_per_s\":%.3f,"
           "\"tflops_equiv\":%.2f,\"x_fp32_filter\":%.2f,\"cand_per_ray\":%.3f}\n",
           name, G, NC, NBUF, (flags & F_MMA) ? 1 : 0, (flags & F_SIGN3) ? 3 : (flags & F_SIGN) ? 1 : 0, best, tests / best * 1e-9, tests * 17.0 / best * 1e-9,
           tests * 17.0 / best * 1e-9 / 65.8, (double)out[0] / rays);
    fflush(stdout);
}

int main(int argc, char** argv)
{
    int iters = argc > 1 ? atoi(argv[1]) : 200;
    cudaDeviceProp prop; CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    printf("{\"probe\":\"device\",\"name\":\"%s\",\"sms\":%d}\n", prop.name, sms);

    // the small spheres of the final scene (main.rs:59-102): a 22 x 22 grid of r = 0.2 plus three r = 1
    std::vector<Sph> sp;
    for (int A = -11; A < 11; ++A)
        for (int B = -11; B < 11; ++B) {
            Sph s{A + 0.9 * urand(), 0.2, B + 0.9 * urand(), 0.2};
            if (std::sqrt((s.x - 4) * (s.x - 4) + (s.z) * (s.z)) > 0.9) sp.push_back(s);
        }
    sp.push_back({0, 1, 0, 1}); sp.push_back({-4, 1, 0, 1}); sp.push_back({4, 1, 0, 1});
    const int n_real = (int)sp.size();
    const int npad = (n_real + 383) / 384 * 384;                   // a multiple of every chunk size tried (32, 64, 96, 128)
    double Rs = 0; for (auto& s : sp) Rs = std::max(Rs, std::sqrt(s.x * s.x + s.y * s.y + s.z * s.z) + s.r);
    const float Rp = std::exp2(std::ceil(std::log2(Rs)));          // power of two >= the bounding radius
    FeatScale sc{Rp, 1.0f, Rp * 0.5f, 1.0f / Rp};
    const double slack = 4e-3 * (Rs * Rs / 256.0);                 // conservative: bounds the split + accumulation error (checked below)

    // B image: features of sphere j, scaled, split hi/lo in fp16, canonical no-swizzle K-major layout
    std::vector<unsigned char> bimg(2 * RT_UMMA_B_BLOCK_BYTES(npad), 0);
    auto put = [&](int j, int k, double val) {
        const float x = (float)val; const unsigned short h = f2h(x); const unsigned short l = f2h(x - h2f(h));
        memcpy(&bimg[b_offset(j, k)], &h, 2);
        memcpy(&bimg[RT_UMMA_B_BLOCK_BYTES(npad) + b_offset(j, k)], &l, 2);
    };
    for (int j = 0; j < npad; ++j) {
        if (j < n_real) {
            const Sph& s = sp[j];
            put(j, 0, (s.r * s.r - (s.x * s.x + s.y * s.y + s.z * s.z) + slack) / sc.s0);
            put(j, 1, s.x / sc.s1); put(j, 2, s.y / sc.s1); put(j, 3, s.z / sc.s1);
            put(j, 4, s.x * s.x / sc.s4); put(j, 5, s.y * s.y / sc.s4); put(j, 6, s.z * s.z / sc.s4);
            put(j, 7, s.x * s.y / sc.s4); put(j, 8, s.x * s.z / sc.s4); put(j, 9, s.y * s.z / sc.s4);
            put(j, 10, 1.0 / sc.s10);
        } else {
            put(j, 0, -4.0 * Rp * Rp / sc.s0);                     // padding: never passes ...
            put(j, 10, 1.0 / sc.s10);                              // ... a dead ray (R_10 = -1, the rest 0) included
        }
    }

    // rays: camera rays from (13,2,3) and bounce rays leaving random points of the sphere field
    const int G_MAX = 6;
    const int n_rays = sms * G_MAX * 128;
    std::vector<float4> ro(n_rays), rd(n_rays);
    for (int i = 0; i < n_rays; ++i) {
        double o[3], d[3];
        if (i % 3 == 0) {
            o[0] = 13 + 0.05 * (urand() - 0.5); o[1] = 2 + 0.05 * (urand() - 0.5); o[2] = 3 + 0.05 * (urand() - 0.5);
            const double tx = 22 * (urand() - 0.5), ty = 3 * urand() - 0.5, tz = 22 * (urand() - 0.5);
            d[0] = tx - o[0]; d[1] = ty - o[1]; d[2] = tz - o[2];
        } else {
            o[0] = 24 * (urand() - 0.5); o[1] = (i % 3 == 1) ? 0.0 : 0.4 * urand(); o[2] = 24 * (urand() - 0.5);
            const double z = (i % 3 == 1) ? urand() : 2 * urand() - 1, ph = 6.283185307179586 * urand(), rr = std::sqrt(std::max(0.0, 1 - z * z));
            d[0] = rr * std::cos(ph); d[1] = z; d[2] = rr * std::sin(ph);
        }
        const double l = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        ro[i] = make_float4((float)o[0], (float)o[1], (float)o[2], 0.f);
        rd[i] = make_float4((float)(d[0] / l), (float)(d[1] / l), (float)(d[2] / l), 0.f);
    }

    ProbeArgs a{};
    unsigned char* d_b; float4 *d_o, *d_d; float* d_dump; unsigned long long* d_out;
    const int dump_rays = 32 * 128;
    CK(cudaMalloc(&d_b, bimg.size())); CK(cudaMemcpy(d_b, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_o, n_rays * sizeof(float4))); CK(cudaMemcpy(d_o, ro.data(), n_rays * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_d, n_rays * sizeof(float4))); CK(cudaMemcpy(d_d, rd.data(), n_rays * sizeof(float4), cudaMemcpyHostToDevice));
    CK(cudaMalloc(&d_dump, (size_t)dump_rays * npad * sizeof(float)));
    CK(cudaMalloc(&d_out, 16));
    long long* d_trace; CK(cudaMalloc(&d_trace, 64 * sizeof(long long))); CK(cudaMemset(d_trace, 0, 64 * sizeof(long long)));
    a.bimg = d_b; a.rays_o = d_o; a.rays_d = d_d; a.npad = npad; a.iters = iters; a.R2 = (float)(Rs * Rs * 1.0001); a.sc = sc;
    a.dump = d_dump; a.dump_rays = dump_rays; a.out = d_out; a.trace = d_trace;
    // > half of the SM's shared memory: exactly one CTA per SM (each CTA allocates all 512 TMEM columns)
    const size_t smem_bytes = std::max<size_t>(2 * RT_UMMA_B_BLOCK_BYTES(npad) + 1024, 120 * 1024);

    // ---- accuracy: both readings of LBO/SBO --------------------------------------------------------------------------
    for (int swap = 0; swap < 3; ++swap) {                        // 2: fp16 accumulator (F_D16)
        ProbeArgs c = a; c.iters = 1;
        CK(cudaMemset(d_dump, 0, (size_t)dump_rays * npad * sizeof(float)));
        CK(cudaMemset(d_out, 0, 16));
        auto k0 = probe_kernel<2, 64, 2, F_MMA | F_SIGN | F_DUMP>; auto k1 = probe_kernel<2, 64, 2, F_MMA | F_SIGN | F_DUMP | F_SWAP>;
        auto k2 = probe_kernel<2, 64, 2, F_MMA | F_SIGN | F_DUMP | F_D16>;
        auto kk = swap == 2 ? k2 : swap ? k1 : k0;
        CK(cudaFuncSetAttribute(kk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        kk<<<sms, 320, smem_bytes>>>(c);
        CK(cudaDeviceSynchronize());
        std::vector<float> D((size_t)dump_rays * npad);
        CK(cudaMemcpy(D.data(), d_dump, D.size() * sizeof(float), cudaMemcpyDeviceToHost));
        double max_err = 0, sum_err = 0, max_err_small = 0, max_excess = 0; long n = 0, false_neg = 0, exact_pos = 0, filt_pos = 0, pad_pos = 0, dead = 0;
        for (int r = 0; r < dump_rays; ++r) {
            const double o[3] = {ro[r].x, ro[r].y, ro[r].z}, d[3] = {rd[r].x, rd[r].y, rd[r].z};
            const double aa = d[0] * d[0] + d[1] * d[1] + d[2] * d[2];
            // is the ray "live" (same decision as the kernel, in f64: rays near the boundary are skipped)
            const double tt = (o[0] * d[0] + o[1] * d[1] + o[2] * d[2]) / aa;
            const double f[3] = {o[0] - tt * d[0], o[1] - tt * d[1], o[2] - tt * d[2]};
            const double ff = f[0] * f[0] + f[1] * f[1] + f[2] * f[2];
            if (std::fabs(ff - a.R2) < 1e-2) continue;
            const bool live = ff < a.R2;
            if (!live) ++dead;
            for (int j = 0; j < npad; ++j) {
                const float got = D[(size_t)r * npad + j];
                if (j >= n_real) { if (!(got < 0)) ++pad_pos; continue; }
                const Sph& s = sp[j];
                const double oc[3] = {s.x - o[0], s.y - o[1], s.z - o[2]};
                const double hb = oc[0] * d[0] + oc[1] * d[1] + oc[2] * d[2];
                const double disc = hb * hb - aa * (oc[0] * oc[0] + oc[1] * oc[1] + oc[2] * oc[2] - s.r * s.r);   // sphere.rs:24 (quarter form)
                if (disc >= 0) ++exact_pos;
                if (!live) { if (disc >= 0) ++false_neg; continue; }       // a dead ray must not be able to hit anything
                const double err = std::fabs((double)got - (disc + slack * aa));
                max_err = std::max(max_err, err); sum_err += err; ++n;
                if (std::fabs(disc + slack * aa) < 0.05) max_err_small = std::max(max_err_small, err);          // where the sign is decided
                max_excess = std::max(max_excess, err - std::fabs(disc + slack * aa) * (1.0 / 1024.0));         // beyond one fp16 rounding of the result
                if (got >= 0) ++filt_pos;
                if (disc >= 0 && !(got >= 0)) ++false_neg;
            }
        }
        printf("{\"probe\":\"umma_accuracy\",\"lbo_sbo_swapped\":%d,\"rays\":%d,\"spheres\":%d,\"npad\":%d,\"scale_Rp\":%.1f,\"slack\":%.3e,\"max_abs_err\":%.4e,"
               "\"mean_abs_err\":%.4e,\"false_negatives\":%ld,\"exact_positive\":%ld,\"filter_positive\":%ld,\"padding_positive\":%ld,\"dead_rays\":%ld,"
               "\"d_fp16\":%d,\"max_abs_err_where_small\":%.4e,\"max_err_beyond_fp16_rounding\":%.4e,\"mask_mismatches\":%llu}\n",
               swap == 1, dump_rays, n_real, npad, Rp, slack, max_err, n ? sum_err / n : 0.0, false_neg, exact_pos, filt_pos, pad_pos, dead,
               swap == 2, max_err_small, max_excess, [&] { unsigned long long o[2]; CK(cudaMemcpy(o, d_out, 16, cudaMemcpyDeviceToHost)); return swap == 2 ? o[1] : 0ull; }());
        fflush(stdout);
    }

    // ---- timeline of one scan (block 0, group 0, warp 0, lane 0): stamps at scan start, after the A store + group barrier, after
    //      issuing chunks 0 and 1, then per chunk: full-wait done, loads done, next chunk issued ------------------------------------
    for (int mode = 0; mode < 2; ++mode) {
        ProbeArgs c = a; c.iters = 8;
        auto k0 = probe_kernel<3, 64, 2, F_MMA | F_SIGN | F_TRACE>; auto k1 = probe_kernel<3, 64, 2, F_MMA | F_SIGN | F_TRACE | F_NOLD>;
        CK(cudaFuncSetAttribute(mode ? k1 : k0, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes));
        (mode ? k1 : k0)<<<sms, 480, smem_bytes>>>(c);
        CK(cudaDeviceSynchronize());
        long long tr[64]; CK(cudaMemcpy(tr, d_trace, sizeof tr, cudaMemcpyDeviceToHost));
        printf("{\"probe\":\"umma_trace\",\"groups\":3,\"chunk\":64,\"noload\":%d,\"cycles_since_scan_start\":[", mode);
        for (int i = 1; i < 64 && tr[i]; ++i) printf("%s%lld", i > 1 ? "," : "", tr[i] - tr[0]);
        printf("]}\n");
    }

    // ---- raw MMA rate (one CTA) ----
    {
        long long* d_rate; CK(cudaMalloc(&d_rate, 16));
        run_rate<32, 1, 1>(d_b, npad, d_rate, smem_bytes); run_rate<48, 1, 1>(d_b, npad, d_rate, smem_bytes);
        run_rate<64, 1, 1>(d_b, npad, d_rate, smem_bytes); run_rate<64, 0, 1>(d_b, npad, d_rate, smem_bytes);
        run_rate<128, 1, 1>(d_b, npad, d_rate, smem_bytes); run_rate<256, 1, 1>(d_b, npad, d_rate, smem_bytes);
        run_rate<96, 1, 1>(d_b, npad, d_rate, smem_bytes); run_rate<64, 1, 0>(d_b, npad, d_rate, smem_bytes); run_rate<256, 1, 0>(d_b, npad, d_rate, smem_bytes);
    }

    // ---- throughput ----
    run<6, 64, 1, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<6, 64, 1, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<6, 32, 2, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<6, 32, 1, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<4, 64, 1, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<3, 64, 2, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<3, 128, 1, F_MMA | F_SIGN | F_D16>("full, fp16 D + PRMT/LOP3 signs", a, n_real, smem_bytes, sms);
    run<4, 64, 1, F_SIGN | F_D16>("ld+sign, fp16 D", a, n_real, smem_bytes, sms);
    run<4, 64, 1, F_MMA | F_D16>("mma+ld, fp16 D", a, n_real, smem_bytes, sms);
    run<3, 64, 2, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<2, 96, 2, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<3, 96, 1, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<4, 64, 1, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<5, 64, 1, F_MMA | F_SIGN>("full", a, n_real, smem_bytes, sms);
    run<3, 96, 1, F_MMA | F_SIGN3>("full, 3-word AND", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_MMA | F_SIGN3>("full, 3-word AND", a, n_real, smem_bytes, sms);
    run<2, 96, 2, F_MMA | F_SIGN3>("full, 3-word AND", a, n_real, smem_bytes, sms);
    run<3, 64, 2, F_MMA | F_NOLD>("mma only", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_MMA | F_NOLD>("mma only", a, n_real, smem_bytes, sms);
    run<3, 64, 2, F_MMA>("mma+ld", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_MMA>("mma+ld", a, n_real, smem_bytes, sms);
    run<3, 64, 2, F_SIGN>("ld+sign", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_SIGN>("ld+sign", a, n_real, smem_bytes, sms);
    run<4, 96, 1, F_SIGN3>("ld+sign, 3-word AND", a, n_real, smem_bytes, sms);
    run<3, 64, 2, 0>("ld", a, n_real, smem_bytes, sms);
    return 0;
}
