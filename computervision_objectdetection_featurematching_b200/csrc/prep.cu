// prep.cu — operand preparation for the tensor-core match kernel.
//
// The match stage (reference src/TestsDetector.cpp:59-60, cv::BFMatcher NORM_L2) is evaluated as the
// dense contraction  acc' = 2 q.t - ||t||^2  so that  d^2 = ||q||^2 - acc'.  SIFT descriptors are
// integers 0..255: they are exact in bf16, every product and every partial sum is
// an integer below 2^24, so the fp32 accumulation in TMEM is exact and no re-rank is needed (App. A.6).
// ||t||^2 is folded into the contraction through KAUG extra K columns: the train row carries -p1, -p2, -p3 with
// p1 = bf16(tn), p2 = bf16(tn - p1), p3 = bf16(tn - p1 - p2) — three 8-bit pieces that add up to the fp32 value of
// tn = ||t||^2 exactly (for integer rows tn < 2^24 is the exact norm and every piece is an integer) — against ones on
// the query side.  *flag collects what the rows are: bit 0 = some value is not an integer in [0, 255] (the
// accumulator is then only an approximation: the call is served by the candidate + fp32 re-rank path), bit 1 = a
// norm is not finite or huge (exact SIMT kernel).  *tnmax_bits (train side) = largest ||t||^2 as float bits, for the
// error bound of the re-rank path.
#include "common.cuh"

namespace cvg {

// one warp per (padded) row; lane handles 4 consecutive floats.  X, Xb, Xlo, Xaug, norms point at row 0 of the matrix
// (segment) the row belongs to.
__device__ __forceinline__ void prep_one_row(const float* __restrict__ X, int n_rows, int row, int is_train,
                                             __nv_bfloat16* __restrict__ Xb, __nv_bfloat16* __restrict__ Xlo,
                                             __nv_bfloat16* __restrict__ Xaug, float* __restrict__ norms,
                                             int* __restrict__ nonint_flag, int* __restrict__ tnmax_bits)
{
    const int lane = threadIdx.x & 31;
    __nv_bfloat162* out = reinterpret_cast<__nv_bfloat162*>(Xb + (size_t)row * DIM) + lane * 2;
    __nv_bfloat16* aug = Xaug + (size_t)row * KAUG;
    __nv_bfloat162* outlo = Xlo ? reinterpret_cast<__nv_bfloat162*>(Xlo + (size_t)row * DIM) + lane * 2 : nullptr;
    if (row >= n_rows) {                      // padding row: never selected / never read back
        out[0] = __floats2bfloat162_rn(0.f, 0.f);
        out[1] = __floats2bfloat162_rn(0.f, 0.f);
        if (outlo) { outlo[0] = __floats2bfloat162_rn(0.f, 0.f); outlo[1] = __floats2bfloat162_rn(0.f, 0.f); }
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
    if (lane == 0) {
        int bits = all_ok ? 0 : 1;
        if (!(ss < 1e30f)) bits |= 2;                        // inf / NaN / out of any sensible range
        if (bits) atomicOr(nonint_flag, bits);
        if (tnmax_bits && ss < 1e30f) atomicMax(tnmax_bits, __float_as_int(ss));
    }
    const float sc = is_train ? 1.f : 2.f;
    out[0] = __floats2bfloat162_rn(e[0] * sc, e[1] * sc);
    out[1] = __floats2bfloat162_rn(e[2] * sc, e[3] * sc);
    if (outlo) {                                             // x = hi + lo + O(2^-18 |x|); lo == 0 for integers 0..255
        float lo[4];
        #pragma unroll
        for (int k = 0; k < 4; k++) {
            const float x = e[k] * sc;                       // exact (sc is 1 or 2)
            const float hi = __bfloat162float(__float2bfloat16(x));
            lo[k] = (fabsf(x) < 1e30f) ? __fsub_rn(x, hi) : 0.f;
        }
        outlo[0] = __floats2bfloat162_rn(lo[0], lo[1]);
        outlo[1] = __floats2bfloat162_rn(lo[2], lo[3]);
    }
    if (lane < KAUG) {
        float a = 0.f;
        if (is_train) {
            const float tn = ss < 1e30f ? ss : 0.f;
            const float p1 = __bfloat162float(__float2bfloat16(tn));
            const float r1 = __fsub_rn(tn, p1);
            const float p2 = __bfloat162float(__float2bfloat16(r1));
            const float p3 = __fsub_rn(r1, p2);                  // <= 8 significant bits left: exact in bf16
            if (lane == 0) a = -p1;
            if (lane == 1) a = -p2;
            if (lane == 2) a = -p3;
        } else {
            a = lane < 3 ? 1.f : 0.f;
        }
        aug[lane] = __float2bfloat16(a);
    }
    if (lane == 0 && norms) norms[row] = ss;
}

// A CTA of 8 warps converts PREP_ROWS_PER_CTA consecutive rows; the row kinds and the largest norm are collected in
// shared memory and reach the global words once per CTA.  (One atomicMax per ROW on one address — 524 288 of them for a
// batch of 64 scenes — serialised in L2 and was the whole cost of the kernel: 0.52 ms for 550 MB of traffic.)
constexpr int PREP_ROWS_PER_CTA = 64;

__global__ void prep_rows_kernel(const float* __restrict__ X, int n_rows, int n_pad, int is_train,
                                 __nv_bfloat16* __restrict__ Xb, __nv_bfloat16* __restrict__ Xlo,
                                 __nv_bfloat16* __restrict__ Xaug,
                                 float* __restrict__ norms, int* __restrict__ nonint_flag, int* __restrict__ tnmax_bits)
{
    __shared__ int s_flag, s_max;
    if (threadIdx.x == 0) { s_flag = 0; s_max = 0; }
    __syncthreads();
    const int wpc = blockDim.x >> 5;
    for (int r = threadIdx.x >> 5; r < PREP_ROWS_PER_CTA; r += wpc) {
        const int64_t row = (int64_t)blockIdx.x * PREP_ROWS_PER_CTA + r;
        if (row < n_pad) prep_one_row(X, n_rows, (int)row, is_train, Xb, Xlo, Xaug, norms, &s_flag, tnmax_bits ? &s_max : nullptr);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_flag) atomicOr(nonint_flag, s_flag);
        if (tnmax_bits && s_max > 0) atomicMax(tnmax_bits, s_max);
    }
}

// All train segments of a scene batch in one launch: padded row r belongs to the segment whose padded range holds it
// (binary search over the segment table; segments are padded to multiples of TILE_N rows).
__global__ void prep_train_segments_kernel(const float* __restrict__ X, const PrepSeg* __restrict__ segs, int n_segs,
                                           int64_t rows_pad_total, __nv_bfloat16* __restrict__ Xb,
                                           __nv_bfloat16* __restrict__ Xlo, __nv_bfloat16* __restrict__ Xaug,
                                           int* __restrict__ nonint_flag, int* __restrict__ tnmax_bits)
{
    // the 64 rows of a CTA lie in one segment (segments are padded to multiples of 256 rows): one search per CTA,
    // by warp 0 with the table probes spread over its lanes
    __shared__ PrepSeg s_seg;
    __shared__ int s_flag, s_max;
    const int64_t prow0 = (int64_t)blockIdx.x * PREP_ROWS_PER_CTA;
    if (prow0 >= rows_pad_total) return;
    if (threadIdx.x == 0) { s_flag = 0; s_max = 0; }
    if (threadIdx.x < 32) {
        int lo = 0, hi = n_segs - 1;                             // last segment with pad_row0 <= prow0
        while (hi - lo >= 32) {
            const int step = (hi - lo + 32) / 32;
            const int probe = min(lo + (int)threadIdx.x * step, hi);
            const unsigned ok = __ballot_sync(0xffffffffu, segs[probe].pad_row0 <= prow0);
            const int last = 31 - __clz(ok);                     // lane 0 probes lo, which always qualifies
            const int nlo = min(lo + last * step, hi);
            hi = min(nlo + step - 1, hi); lo = nlo;
        }
        const int probe = min(lo + (int)threadIdx.x, hi);
        const unsigned ok = __ballot_sync(0xffffffffu, segs[probe].pad_row0 <= prow0);
        const int idx = min(lo + (31 - __clz(ok)), hi);
        if (threadIdx.x == 0) s_seg = segs[idx];
    }
    __syncthreads();
    const PrepSeg g = s_seg;
    const int wpc = blockDim.x >> 5;
    for (int r = threadIdx.x >> 5; r < PREP_ROWS_PER_CTA; r += wpc) {
        const int64_t prow = prow0 + r;
        if (prow < rows_pad_total)
            prep_one_row(X + g.f32_row0 * DIM, g.rows, (int)(prow - g.pad_row0), 1, Xb + g.pad_row0 * DIM,
                         Xlo ? Xlo + g.pad_row0 * DIM : nullptr, Xaug + g.pad_row0 * KAUG, nullptr, &s_flag,
                         tnmax_bits ? &s_max : nullptr);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_flag) atomicOr(nonint_flag, s_flag);
        if (tnmax_bits && s_max > 0) atomicMax(tnmax_bits, s_max);
    }
}

void launch_prep_train_segments(const float* X, const PrepSeg* segs_dev, int n_segs, int64_t rows_pad_total,
                                __nv_bfloat16* Xb, __nv_bfloat16* Xlo, __nv_bfloat16* Xaug, int* nonint_flag,
                                int* tnmax_bits, cudaStream_t st)
{
    if (rows_pad_total <= 0 || n_segs <= 0) return;
    const int threads = 256;
    const unsigned blocks = (unsigned)((rows_pad_total + PREP_ROWS_PER_CTA - 1) / PREP_ROWS_PER_CTA);
    prep_train_segments_kernel<<<blocks, threads, 0, st>>>(X, segs_dev, n_segs, rows_pad_total, Xb, Xlo, Xaug, nonint_flag,
                                                           tnmax_bits);
}

void launch_prep_rows(const float* X, int n_rows, int n_pad, int is_train, __nv_bfloat16* Xb, __nv_bfloat16* Xlo,
                      __nv_bfloat16* Xaug, float* norms, int* nonint_flag, int* tnmax_bits, cudaStream_t st)
{
    if (n_pad <= 0) return;
    const int threads = 256;
    const int blocks = (int)(((int64_t)n_pad + PREP_ROWS_PER_CTA - 1) / PREP_ROWS_PER_CTA);
    prep_rows_kernel<<<blocks, threads, 0, st>>>(X, n_rows, n_pad, is_train, Xb, Xlo, Xaug, norms, nonint_flag, tnmax_bits);
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
