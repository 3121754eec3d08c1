// scene_gen.cpp — seeded scene builder: random_scene of
// /root/reference/src/main.rs:59-102 with an explicit seed, so that the oracle and the GPU render
// the same world.  Host-only (no CUDA calls); the draw ORDER follows the reference line by line,
// the random stream itself is Philox4x32-10 because thread_rng (main.rs:60) cannot be seeded.
#include "../../include/rtiow_cuda.h"
#include <cmath>
#include <cstdint>

namespace {

struct Stream {
    uint64_t seed; uint64_t k = 0;
    static void philox(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t o[4])
    {
        for (int r = 0; r < 10; ++r) {
            uint64_t p0 = (uint64_t)0xD2511F53u * c0, p1 = (uint64_t)0xCD9E8D57u * c2;
            uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c3 ^ k1;
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
            k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
        }
        o[0] = c0; o[1] = c1; o[2] = c2; o[3] = c3;
    }
    double gen()                                  // rng.gen::<f64>(): 53-bit uniform in [0,1)
    {
        uint32_t o[4];
        philox((uint32_t)k, (uint32_t)(k >> 32), 0u, 0x5343454Eu /* "SCEN" */, (uint32_t)seed, (uint32_t)(seed >> 32), o);
        ++k;
        const uint64_t bits = ((uint64_t)o[1] << 32) | o[0];
        return (double)(bits >> 11) * (1.0 / 9007199254740992.0);
    }
    double range(double lo, double hi) { return lo + (hi - lo) * gen(); }   // gen_range(lo..hi) / (lo..=hi)
};

struct Out {
    uint32_t cap, n = 0; bool overflow = false;
    double *cx, *cy, *cz, *radius, *albedo, *param; uint32_t* kind;
    void push(double x, double y, double z, double r, uint32_t k, double ar, double ag, double ab, double prm)   // world.push(Box::new(Sphere::new(..)))
    {
        if (n >= cap) { overflow = true; return; }
        cx[n] = x; cy[n] = y; cz[n] = z; radius[n] = r; kind[n] = k;
        albedo[3 * n] = ar; albedo[3 * n + 1] = ag; albedo[3 * n + 2] = ab; param[n] = prm;
        ++n;
    }
};

}  // namespace

extern "C" int rtiow_random_scene(uint64_t seed, int32_t half_extent, int32_t material_mode, uint32_t cap, double* cx, double* cy, double* cz,
                                  double* radius, uint32_t* mat_kind, double* albedo_rgb, double* mat_param, uint32_t* out_n)
{
    if (!cx || !cy || !cz || !radius || !mat_kind || !albedo_rgb || !mat_param || !out_n) return RTIOW_ERR_INVALID_ARG;
    if (half_extent < 0 || half_extent > 2000 || material_mode < 0 || material_mode > 3) return RTIOW_ERR_INVALID_ARG;
    Stream rng{ seed };
    Out w{ cap, 0, false, cx, cy, cz, radius, albedo_rgb, mat_param, mat_kind };

    w.push(0, -1000, 0, 1000, RTIOW_MAT_LAMBERTIAN, 0.5, 0.5, 0.5, 0);                       // main.rs:63-64

    for (int a = -half_extent; a <= half_extent; ++a) {                                      // main.rs:66
        for (int b = -half_extent; b <= half_extent; ++b) {                                  // main.rs:67
            const double a_prime = (double)a + 0.9 * rng.gen();                              // main.rs:68
            const double b_prime = (double)b + 0.9 * rng.gen();                              // main.rs:69
            const double dx = a_prime - 4.0, dy = 0.2 - 0.2, dz = b_prime - 0.0;
            if (std::sqrt(dx * dx + dy * dy + dz * dz) > 0.9) {                              // main.rs:72
                const double choose = rng.gen();                                             // main.rs:73
                int branch = choose <= 0.8 ? 0 : (choose <= 0.95 ? 1 : 2);                   // main.rs:74,78,83 (tested in order)
                if (material_mode != 0) branch = material_mode - 1;
                if (branch == 0) {
                    const double r1 = rng.gen(), g1 = rng.gen(), b1 = rng.gen();             // Color::random() (vec3.rs:21-24)
                    const double r2 = rng.gen(), g2 = rng.gen(), b2 = rng.gen();             // * Color::random()  (main.rs:75)
                    w.push(a_prime, 0.2, b_prime, 0.2, RTIOW_MAT_LAMBERTIAN, r1 * r2, g1 * g2, b1 * b2, 0);
                } else if (branch == 1) {
                    const double r = rng.range(0.5, 1), g = rng.range(0.5, 1), bb = rng.range(0.5, 1);   // main.rs:79
                    const double fuzz = rng.range(0.0, 0.5);                                 // main.rs:80
                    w.push(a_prime, 0.2, b_prime, 0.2, RTIOW_MAT_METAL, r, g, bb, fuzz);
                } else {
                    w.push(a_prime, 0.2, b_prime, 0.2, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);  // main.rs:84
                }
            }
        }
    }
    switch (material_mode) {
    case 0:                                                                                   // main.rs:92-99
        w.push(0, 1, 0, 1.0, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);
        w.push(-4, 1, 0, 1.0, RTIOW_MAT_LAMBERTIAN, 0.4, 0.2, 0.1, 0);
        w.push(4, 1, 0, 1.0, RTIOW_MAT_METAL, 0.7, 0.6, 0.5, 0.0);
        break;
    case 1:
        w.push(0, 1, 0, 1.0, RTIOW_MAT_LAMBERTIAN, 0.6, 0.6, 0.6, 0);
        w.push(-4, 1, 0, 1.0, RTIOW_MAT_LAMBERTIAN, 0.4, 0.2, 0.1, 0);
        w.push(4, 1, 0, 1.0, RTIOW_MAT_LAMBERTIAN, 0.7, 0.6, 0.5, 0);
        break;
    case 2:
        w.push(0, 1, 0, 1.0, RTIOW_MAT_METAL, 0.8, 0.8, 0.8, 0.1);
        w.push(-4, 1, 0, 1.0, RTIOW_MAT_METAL, 0.4, 0.2, 0.1, 0.3);
        w.push(4, 1, 0, 1.0, RTIOW_MAT_METAL, 0.7, 0.6, 0.5, 0.0);
        break;
    default:
        w.push(0, 1, 0, 1.0, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);
        w.push(0, 1, 0, -0.9, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);      // hollow shell: negative radius (sphere.rs:45-51)
        w.push(-4, 1, 0, 1.0, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);
        w.push(4, 1, 0, 1.0, RTIOW_MAT_DIELECTRIC, 1, 1, 1, 1.5);
        break;
    }
    *out_n = w.n;
    return w.overflow ? RTIOW_ERR_NOMEM : RTIOW_OK;
}

// ------------------------------------------------------------------------------------------------
// Scene dump / load (second half): one world, many renderers.  The reference builds its world in code
// (main.rs:59-102) from an unseedable thread_rng; a text file of the same Sphere / material records lets the oracle, this
// library and a real `cargo` build of the reference (INTEGRATION.md shows the 20-line Rust reader) render the SAME spheres.
//   rtiow-scene 1
//   n <count>
//   <cx> <cy> <cz> <radius> <kind> <albedo_r> <albedo_g> <albedo_b> <param>      one line per sphere, list order
// kind: 0 Lambertian, 1 Metal (param = fuzz), 2 Dialectric (param = ir).  Numbers are printed with 17 significant digits, so
// every f64 survives the round trip bit for bit.
// ------------------------------------------------------------------------------------------------
#include <cstdio>
#include <cinttypes>

extern "C" int rtiow_scene_save(const char* path, uint32_t n, const double* cx, const double* cy, const double* cz, const double* radius,
                                const uint32_t* mat_kind, const double* albedo_rgb, const double* mat_param)
{
    if (!path || (n > 0 && (!cx || !cy || !cz || !radius || !mat_kind || !albedo_rgb || !mat_param))) return RTIOW_ERR_INVALID_ARG;
    for (uint32_t i = 0; i < n; ++i) if (mat_kind[i] > RTIOW_MAT_DIELECTRIC) return RTIOW_ERR_UNSUPPORTED;
    FILE* f = std::fopen(path, "w");
    if (!f) return RTIOW_ERR_INVALID_ARG;
    std::fprintf(f, "rtiow-scene 1\nn %" PRIu32 "\n", n);
    for (uint32_t i = 0; i < n; ++i)
        std::fprintf(f, "%.17g %.17g %.17g %.17g %" PRIu32 " %.17g %.17g %.17g %.17g\n", cx[i], cy[i], cz[i], radius[i], mat_kind[i],
                     albedo_rgb[3 * i], albedo_rgb[3 * i + 1], albedo_rgb[3 * i + 2], mat_param[i]);
    const bool bad = std::ferror(f) != 0;
    return (std::fclose(f) != 0 || bad) ? RTIOW_ERR_INVALID_ARG : RTIOW_OK;
}

// cap = 0 (arrays may be NULL): only the count is returned in *out_n.  More spheres in the file than cap: RTIOW_ERR_NOMEM,
// *out_n = the count needed.
extern "C" int rtiow_scene_load(const char* path, uint32_t cap, double* cx, double* cy, double* cz, double* radius, uint32_t* mat_kind,
                                double* albedo_rgb, double* mat_param, uint32_t* out_n)
{
    if (!path || !out_n) return RTIOW_ERR_INVALID_ARG;
    if (cap > 0 && (!cx || !cy || !cz || !radius || !mat_kind || !albedo_rgb || !mat_param)) return RTIOW_ERR_INVALID_ARG;
    FILE* f = std::fopen(path, "r");
    if (!f) return RTIOW_ERR_INVALID_ARG;
    int version = 0; uint32_t n = 0;
    if (std::fscanf(f, " rtiow-scene %d n %" SCNu32, &version, &n) != 2 || version != 1) { std::fclose(f); return RTIOW_ERR_UNSUPPORTED; }
    *out_n = n;
    if (cap == 0) { std::fclose(f); return RTIOW_OK; }
    if (n > cap) { std::fclose(f); return RTIOW_ERR_NOMEM; }
    for (uint32_t i = 0; i < n; ++i) {
        double v[8]; uint32_t k = 0;
        if (std::fscanf(f, "%lf %lf %lf %lf %" SCNu32 " %lf %lf %lf %lf", &v[0], &v[1], &v[2], &v[3], &k, &v[4], &v[5], &v[6], &v[7]) != 9 ||
            k > RTIOW_MAT_DIELECTRIC) { std::fclose(f); return RTIOW_ERR_INVALID_ARG; }
        cx[i] = v[0]; cy[i] = v[1]; cz[i] = v[2]; radius[i] = v[3]; mat_kind[i] = k;
        albedo_rgb[3 * i] = v[4]; albedo_rgb[3 * i + 1] = v[5]; albedo_rgb[3 * i + 2] = v[6]; mat_param[i] = v[7];
    }
    std::fclose(f);
    return RTIOW_OK;
}
