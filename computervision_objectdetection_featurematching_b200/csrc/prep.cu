// prep.cu — operand preparation for the tensor-core match kernel.
//
// The match stage (reference src/TestsDetector.cpp:59-60, cv::BFMatcher NORM_L2) is evaluated as the
// dense contraction  acc' = 2 q.t - ||t||^2  so that  d^2 = ||q||^2 - acc'.  SIFT descriptors are
// integers 0..255: they are exact in bf16, every product and every partial sum is
// an integer below 2^24, so the fp32 accumulation in TMEM is exact and no re-rank is needed (App. A.6).
// ||t||^2 (< 2^24) is folded into the contraction through KAUG extra K columns: the train row carries
// -(a*65536), -(b*256), -c with ||t||^2 = a*65536 + b*256 + c (each piece exact in bf16) against ones
// on the query side.  Rows that are not integer-valued raise *nonint_flag and the exact fp32 kernel
// serves the call instead.
#include "common.cuh"

namespace cvg {

// one warp per row; lane handles 4 consecutive floats
__global__ void prep_rows_kernel(const float* __restrict__ X, int n_rows, int n_pad, int is_train,
                                 __nv_bfloat16* __restrict__ Xb, __nv_bfloat16* __restrict__ Xaug,
                                 float* __restrict__ norms, int* __restrict__ nonint_flag)
{
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (warp >= n_pad) return;
    const int row = warp;
    __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(Xb + (size_t)row * DIM) + lane * 2;
    __nv_bfloat16* aug = Xaug + (size_t)row * KAUG;
    if (row >= n_rows) {                      // padding row: never selected / never read back
        out[0] = __floats2bfloat162_rn(0.f, 0.f);
        out[1] = __floats2bfloat162_rn(0.f, 0.f);
        if (lane < KAUG) aug[lane] = __float2bfloat16((is_train && lane == 0) ? -1073741824.f : 0.f);
        if (lane == 0 && norms) norms[row] = 0.f;
        return;
    }
    const float4 v = reinterpret_cast<const float4*>(X + (size_t)row * DIM)[lane];
    const float e[4] = { v.x, v.y, v.z, v.w };
    bool ok = true;
    float ss = 0.f;
    #pragma unroll
    for (int k = 0; k < 4; k++) {
        ok = ok && (e[k] >= 0.f) && (e[k] <= 255.f) && (e[k] == floorf(e[k]));
        ss = __fadd_rn(ss, __fmul_rn(e[k], e[k]));
    }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) ss = __fadd_rn(ss, __shfl_xor_sync(0xffffffffu, ss, o));
    const bool all_ok = __all_sync(0xffffffffu, ok);
    if (!all_ok && lane == 0) atomicOr(nonint_flag, 1);
    const float sc = is_train ? 1.f : 2.f;
    out[0] = __floats2bfloat162_rn(e[0] * sc, e[1] * sc);
    out[1] = __floats2bfloat162_rn(e[2] * sc, e[3] * sc);
    if (lane < KAUG) {
        float a = 0.f;
        if (is_train) {
            const uint32_t tn = all_ok ? (uint32_t)ss : 0u;      // exact integer < 2^24 when all_ok
            if (lane == 0) a = -(float)(tn & 0xFF0000u);
            if (lane == 1) a = -(float)(tn & 0x00FF00u);
            if (lane == 2) a = -(float)(tn & 0x0000FFu);
        } else {
            a = lane < 3 ? 1.f : 0.f;
        }
        aug[lane] = __float2bfloat16(a);
    }
    if (lane == 0 && norms) norms[row] = ss;
}

void launch_prep_rows(const float* X, int n_rows, int n_pad, int is_train, __nv_bfloat16* Xb,
                      __nv_bfloat16* Xaug, float* norms, int* nonint_flag, cudaStream_t st)
{
    if (n_pad <= 0) return;
    const int threads = 256;
    const int blocks = (int)(((int64_t)n_pad * 32 + threads - 1) / threads);
    prep_rows_kernel<<<blocks, threads, 0, st>>>(X, n_rows, n_pad, is_train, Xb, Xaug, norms, nonint_flag);
}

__global__ void pack_points_kernel(const float2* __restrict__ src, const float2* __restrict__ dst, int64_t n,
                                   float4* __restrict__ pts)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        float2 a = src[i], b = dst[i];
        pts[i] = make_float4(a.x, a.y, b.x, b.y);
    }
}

void launch_pack_points(const float* src_xy, const float* dst_xy, int64_t n, float4* pts, cudaStream_t st)
{
    if (n <= 0) return;
    pack_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float2*>(src_xy), reinterpret_cast<const float2*>(dst_xy), n, pts);
}

}  // namespace cvg
