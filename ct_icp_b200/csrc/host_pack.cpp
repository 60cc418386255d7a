// host_pack.cpp — the vectorised inner loop of RegisterFrame's only O(N) host pass (Engine::PackAndUpload / PackScan).
// A plain C++ translation unit: nvcc's front end does not declare the AVX intrinsics.
//
// The common layout — contiguous float64 x, y, z and contiguous float64 timestamps (numpy's default) — four points per
// iteration: three 4-double loads, packed double -> float conversions, the float32-representability test as three vector
// compares, two 32-byte non-temporal stores. 1.3-1.7 ns per point on one core against 2.4 ns for the scalar loop; the pass
// is the largest item of the end-to-end call (DESIGN.md §5.2 "Host path").
#include <immintrin.h>

#include <cstddef>
#include <cstdint>

struct float4;   // CUDA's: 16 bytes, (x, y, z, w)

namespace cticp {

bool HostPackHasAvx2() { return __builtin_cpu_supports("avx2"); }

__attribute__((target("avx2"))) void PackBlockF64Avx2(const double *xyz, const double *t, size_t b, size_t e, double mn,
                                                      double inv, bool spans, float4 *dst4, bool *any_lo) {
    float *dst = reinterpret_cast<float *>(dst4);
    const __m256d vmn = _mm256_set1_pd(mn), vinv = _mm256_set1_pd(inv);
    __m256d bad = _mm256_setzero_pd();
    size_t i = b;
    for (; i + 4 <= e; i += 4) {
        const __m256d d0 = _mm256_loadu_pd(xyz + 3 * i), d1 = _mm256_loadu_pd(xyz + 3 * i + 4), d2 = _mm256_loadu_pd(xyz + 3 * i + 8);
        const __m128 a0 = _mm256_cvtpd_ps(d0), a1 = _mm256_cvtpd_ps(d1), a2 = _mm256_cvtpd_ps(d2);   // x0 y0 z0 x1 | y1 z1 x2 y2 | z2 x3 y3 z3
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtps_pd(a0), d0, _CMP_NEQ_UQ));
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtps_pd(a1), d1, _CMP_NEQ_UQ));
        bad = _mm256_or_pd(bad, _mm256_cmp_pd(_mm256_cvtps_pd(a2), d2, _CMP_NEQ_UQ));
        const __m128 al = spans ? _mm256_cvtpd_ps(_mm256_mul_pd(_mm256_sub_pd(_mm256_loadu_pd(t + i), vmn), vinv)) : _mm_set1_ps(1.0f);
        const __m128 p0 = _mm_insert_ps(a0, al, 0x30);                     // x0 y0 z0 | alpha0
        __m128 p1 = _mm_shuffle_ps(a0, a1, _MM_SHUFFLE(1, 0, 3, 3));       // x1 x1 y1 z1
        p1 = _mm_shuffle_ps(p1, p1, _MM_SHUFFLE(3, 3, 2, 0));              // x1 y1 z1 z1
        p1 = _mm_insert_ps(p1, al, 0x70);
        __m128 p2 = _mm_shuffle_ps(a1, a2, _MM_SHUFFLE(0, 0, 3, 2));       // x2 y2 z2 z2
        p2 = _mm_insert_ps(p2, al, 0xB0);
        __m128 p3 = _mm_shuffle_ps(a2, a2, _MM_SHUFFLE(3, 3, 2, 1));       // x3 y3 z3 z3
        p3 = _mm_insert_ps(p3, al, 0xF0);
        _mm256_stream_ps(dst + 4 * i, _mm256_set_m128(p1, p0));        // (b is a multiple of four points and the staging
        _mm256_stream_ps(dst + 4 * i + 8, _mm256_set_m128(p3, p2));    // buffer page-aligned: 32-byte aligned stores)
    }
    if (_mm256_movemask_pd(bad)) *any_lo = true;
    for (; i < e; ++i) {
        const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
        const double a = spans ? (t[i] - mn) * inv : 1.0;
        const float fx = (float) x, fy = (float) y, fz = (float) z;
        _mm_stream_ps(dst + 4 * i, _mm_set_ps((float) a, fz, fy, fx));
        if ((double) fx != x || (double) fy != y || (double) fz != z) *any_lo = true;
    }
    _mm_sfence();
}

}  // namespace cticp
