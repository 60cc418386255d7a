// Host build of the engine's math header (csrc/se3.cuh, csrc/device_map.cuh are host + device): the polynomial replacements
// of libm calls and the reciprocal voxel coordinate against their exact definitions. Plus host models of the device-only
// conversion tricks (float -> double by bit manipulation, int <-> double by magic adds).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <random>

#include "se3.cuh"
#include "device_map.cuh"

using namespace cticp;

static double f32_to_f64_model(float f) {   // se3.cuh f32_to_f64
    uint32_t u;
    memcpy(&u, &f, 4);
    const uint32_t mag = u & 0x7fffffffu;
    const bool tiny = mag < 0x00800000u;
    const uint64_t hi = (u & 0x80000000u) | (tiny ? 0u : (mag >> 3) + 0x38000000u);
    const uint64_t lo = tiny ? 0u : (uint32_t) (u << 29);
    const uint64_t bits = (hi << 32) | lo;
    double d;
    memcpy(&d, &bits, 8);
    return d;
}
static double i32_to_f64_model(int i) {   // se3.cuh i32_to_f64
    const uint64_t bits = ((uint64_t) 0x43300000u << 32) | (uint32_t) (i ^ (int) 0x80000000);
    double d;
    memcpy(&d, &bits, 8);
    return d - 4503601774854144.0;
}

int main() {
    std::mt19937_64 rng(7);
    std::uniform_real_distribution<double> half(-0.5, 0.5), big(-3000.0, 3000.0);
    int bad = 0;
    double worst_s = 0, worst_c = 0;
    for (int i = 0; i < 2000000; ++i) {
        const double x = half(rng);
        const double es = fabs(sin_upto_half(x) - sin(x)), ec = fabs(cos_upto_half(x) - cos(x));
        worst_s = fmax(worst_s, es / fmax(fabs(sin(x)), 1e-300));
        worst_c = fmax(worst_c, ec);
    }
    if (!(worst_s < 4.5e-16 && worst_c < 2.3e-16)) { printf("polynomial error: sin %.3e (rel) cos %.3e\n", worst_s, worst_c); ++bad; }
    // slerp with the hoisted constants against the plain definition, small and large angles
    for (int i = 0; i < 20000; ++i) {
        const double ang = i < 10000 ? 0.2 * fabs(half(rng)) : 2.0 * fabs(half(rng)) + 0.6;
        const Q4 a = qnormalized(Q4{half(rng), half(rng), half(rng), 1.0});
        const Q4 r = qnormalized(Q4{sin(ang / 2) * 0.6, sin(ang / 2) * 0.0, sin(ang / 2) * 0.8, cos(ang / 2)});
        const Q4 b = qmul(a, r);
        const double t = fabs(half(rng)) * 2.0;
        const Q4 s0 = qslerp(a, b, t), s1 = qslerp_c(a, b, t, slerp_consts(a, b));
        const double d = fabs(s0.x - s1.x) + fabs(s0.y - s1.y) + fabs(s0.z - s1.z) + fabs(s0.w - s1.w);
        if (!(d < 4e-15)) { if (bad < 5) printf("slerp mismatch %.3e at angle %.3f\n", d, ang); ++bad; }
    }
    // int(p / res) from the reciprocal: random values, exact multiples, neighbours of multiples
    const double ress[] = {1.0, 0.5, 0.2, 0.25, 0.8, 1.5, 0.1, 3.0};
    for (double res : ress) {
        const double inv = 1.0 / res;
        for (int i = 0; i < 300000; ++i) {
            double p = big(rng);
            if (i % 3 == 0) p = res * (double) (int) (p / res);                       // an exact multiple
            if (i % 9 == 0) p = nextafter(p, i % 2 ? 1e9 : -1e9);                      // one ulp beside it
            if (voxel_coord_rcp(p, res, inv) != voxel_coord(p, res)) { if (bad < 5) printf("voxel_coord_rcp(%.17g, %g)\n", p, res); ++bad; }
        }
    }
    // conversion models
    std::uniform_int_distribution<uint32_t> bits;
    for (int i = 0; i < 3000000; ++i) {
        uint32_t u = bits(rng);
        float f;
        memcpy(&f, &u, 4);
        if (!std::isfinite(f)) continue;
        const double want = fabsf(f) < 1.17549435e-38f ? (std::signbit(f) ? -0.0 : 0.0) : (double) f;   // denormals flush
        const double got = f32_to_f64_model(f);
        if (memcmp(&want, &got, 8) != 0) { if (bad < 5) printf("f32_to_f64(%a)\n", f); ++bad; }
        const int k = (int) u;
        if (i32_to_f64_model(k) != (double) k) { if (bad < 5) printf("i32_to_f64(%d)\n", k); ++bad; }
    }
    if (bad) { printf("FAILED: %d\n", bad); return 1; }
    printf("SE3 MATH OK (sin rel err %.2e, cos abs err %.2e)\n", worst_s, worst_c);
    return 0;
}
