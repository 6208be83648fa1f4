// Host build of csrc/homography_math.cuh for CPU-side unit tests (tests/test_hostmath.py): the very
// functions the CUDA kernels call, compiled with g++ -ffp-contract=off, checked against the oracle.
#include "../../computervision_objectdetection_featurematching_b200/csrc/homography_math.cuh"
#include "../../computervision_objectdetection_featurematching_b200/csrc/jacobi_thread.cuh"
#include <vector>
#include <cstring>

extern "C" {

int hm_run_kernel(const float* src, const float* dst, int n, double* H)
{
    std::vector<float4> p(n);
    for (int i = 0; i < n; i++) p[i] = float4{ src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1] };
    double LtL[81], V[81];
    cvg::PtsStrided P{ p.data(), nullptr };
    return cvg::run_kernel_seq(P, n, H, LtL, V) ? 1 : 0;
}

// element-major (strided) storage of LtL / V, as the hypothesis kernel uses in shared memory
int hm_run_kernel_strided(const float* src, const float* dst, int n, double* H)
{
    std::vector<float4> p(n);
    for (int i = 0; i < n; i++) p[i] = float4{ src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1] };
    constexpr int S = 64;
    std::vector<double> buf(162 * S + 7, -777.0);
    cvg::PtsStrided P{ p.data(), nullptr };
    return cvg::run_kernel_seq<cvg::PtsStrided, S>(P, n, H, buf.data() + 5, buf.data() + 5 + 81 * S) ? 1 : 0;
}

int hm_check_subset4(const float* ms1, const float* ms2) { return cvg::check_subset4(ms1, ms2) ? 1 : 0; }

float hm_reproj_err(const double* H, float Mx, float My, float mx, float my)
{
    float Hf[8];
    for (int i = 0; i < 8; i++) Hf[i] = (float)H[i];
    return cvg::reproj_err(Hf, Mx, My, mx, my);
}

int hm_update_num_iters(double p, double ep, int max_iters) { return cvg::update_num_iters(p, ep, max_iters); }

int hm_lm_refine(const float* src, const float* dst, int n, double* H, int max_iters)
{
    std::vector<float4> p(n);
    for (int i = 0; i < n; i++) p[i] = float4{ src[2 * i], src[2 * i + 1], dst[2 * i], dst[2 * i + 1] };
    cvg::PtsStrided P{ p.data(), nullptr };
    cvg::LmState st;
    double a[81], v2[81];
    return cvg::lm_refine_seq(P, n, H, max_iters, st, a, v2);
}

void hm_jacobi9(const double* A, double* W, double* V)
{
    double a[81];
    memcpy(a, A, sizeof a);
    cvg::jacobi<9>(a, W, V);
}

// thread-per-hypothesis runKernel (csrc/jacobi_thread.cuh) on a 4-point sample, element-major storage
int hm_run_kernel4_thread(const float* src, const float* dst, double* H)
{
    constexpr int S = 32;
    std::vector<double> buf((size_t)cvg::JT_DOUBLES * S + 9, -777.0);
    return cvg::run_kernel4_thread<S>(src, dst, H, buf.data() + 3, true, [](bool x) { return x; }) ? 1 : 0;
}

double hm_det3(const double* H) { return cvg::det3(H); }

uint32_t hm_rng_next(uint64_t* s) { return cvg::rng_next(*s); }

}
