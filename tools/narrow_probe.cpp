#include <immintrin.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <chrono>
#include <thread>
#include <vector>
#include <cstdint>
__attribute__((target("avx2"))) static bool narrow_avx2(const float* src, uint8_t* dst, size_t n)
{
    __m256i bad = _mm256_setzero_si256();
    const __m256i perm = _mm256_setr_epi32(0, 4, 1, 5, 2, 6, 3, 7);
    size_t i = 0;
    for (; i + 32 <= n; i += 32) {
        __m256 a = _mm256_loadu_ps(src + i), b = _mm256_loadu_ps(src + i + 8), c = _mm256_loadu_ps(src + i + 16), d = _mm256_loadu_ps(src + i + 24);
        __m256i ia = _mm256_cvttps_epi32(a), ib = _mm256_cvttps_epi32(b), ic = _mm256_cvttps_epi32(c), id = _mm256_cvttps_epi32(d);
        // exact integers only: float(int(x)) == x, and 0 <= int <= 255 (checked through the saturating packs below)
        __m256 ea = _mm256_cmp_ps(_mm256_cvtepi32_ps(ia), a, _CMP_NEQ_UQ), eb = _mm256_cmp_ps(_mm256_cvtepi32_ps(ib), b, _CMP_NEQ_UQ);
        __m256 ec = _mm256_cmp_ps(_mm256_cvtepi32_ps(ic), c, _CMP_NEQ_UQ), ed = _mm256_cmp_ps(_mm256_cvtepi32_ps(id), d, _CMP_NEQ_UQ);
        bad = _mm256_or_si256(bad, _mm256_castps_si256(_mm256_or_ps(_mm256_or_ps(ea, eb), _mm256_or_ps(ec, ed))));
        __m256i range = _mm256_or_si256(_mm256_or_si256(ia, ib), _mm256_or_si256(ic, id));      // any bit above 255 or the sign bit set?
        bad = _mm256_or_si256(bad, _mm256_andnot_si256(_mm256_set1_epi32(255), range));
        __m256i ab = _mm256_packus_epi32(ia, ib), cd = _mm256_packus_epi32(ic, id);
        __m256i q = _mm256_packus_epi16(ab, cd);
        q = _mm256_permutevar8x32_epi32(q, perm);
        _mm256_storeu_si256((__m256i*)(dst + i), q);
    }
    bool ok = _mm256_testz_si256(bad, bad);
    for (; i < n; i++) { const float x = src[i]; const int v = (int)x; if ((float)v != x || v < 0 || v > 255) ok = false; dst[i] = (uint8_t)v; }
    return ok;
}
int main(int argc, char** argv)
{
    const int T = argc > 1 ? atoi(argv[1]) : 8;
    const size_t n = 68157440;   // 272 MB of floats
    std::vector<float> src(n); std::vector<uint8_t> dst(n), ref(n);
    for (size_t i = 0; i < n; i++) { src[i] = (float)((i * 2654435761u >> 7) & 255); ref[i] = (uint8_t)src[i]; }
    for (int rep = 0; rep < 3; rep++) {
        auto t0 = std::chrono::steady_clock::now();
        std::vector<std::thread> th; std::vector<char> ok(T, 1);
        for (int t = 0; t < T; t++) th.emplace_back([&, t] { size_t a = n * t / T / 32 * 32, b = t == T - 1 ? n : n * (t + 1) / T / 32 * 32; ok[t] = narrow_avx2(src.data() + a, dst.data() + a, b - a); });
        for (auto& x : th) x.join();
        double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        printf("threads %d: %.2f ms  %.1f GB/s read  ok=%d  equal=%d\n", T, s * 1e3, n * 4 / s / 1e9, (int)ok[0], (int)(memcmp(dst.data(), ref.data(), n) == 0));
    }
    src[12345] = 3.5f; printf("noninteger detected: %d\n", !narrow_avx2(src.data(), dst.data(), n));
    src[12345] = 256.f; printf("range detected: %d\n", !narrow_avx2(src.data(), dst.data(), n));
    src[12345] = -1.f; printf("negative detected: %d\n", !narrow_avx2(src.data(), dst.data(), n));
}
