// match_exact.cu — exact fp32 match kernel (general descriptors) and the top-2 merge kernels.
//
// Reference: BFMatcher(NORM_L2).knnMatch(q, t, 2), src/TestsDetector.cpp:59-60.  This kernel
// reproduces cv::batchDistance's fp32 arithmetic bit for bit: 16 partial sums
// acc[a][l] with a = (j/4)%4, l = j%4, separate multiply and add, combine ((a0+a1)+a2)+a3 per lane
// then (v0+v2)+(v1+v3), sqrtf, strict < insertion scanning train rows in ascending order (App. A.2).
// It serves calls whose descriptors are not integer-valued; integer (SIFT) data takes match_tc.cu.
#include "common.cuh"
#include <float.h>

namespace cvg {

constexpr int EX_THREADS = 128;
constexpr int EX_TROWS = 16;          // train rows staged per smem tile

__global__ void __launch_bounds__(EX_THREADS)
match_exact_kernel(const float* __restrict__ Q, int n_query, const float* __restrict__ T,
                   const MatchUnit* __restrict__ units, Top2* __restrict__ parts,
                   const int* __restrict__ run_if_flag)
{
    if (run_if_flag && *run_if_flag == 0) return;
    __shared__ float4 tile[EX_TROWS][DIM / 4];
    const MatchUnit u = units[blockIdx.x];
    const int qrow = u.q_row0 + threadIdx.x;
    const bool qvalid = qrow < n_query;
    float4 q[DIM / 4];
    {
        const float4* qp = reinterpret_cast<const float4*>(Q + (size_t)(qvalid ? qrow : 0) * DIM);
        #pragma unroll
        for (int j = 0; j < DIM / 4; j++) q[j] = qp[j];
    }
    int n_cols = u.n_tiles * TILE_N;
    if (n_cols > u.seg_cols - u.t_local0) n_cols = u.seg_cols - u.t_local0;
    float bd0 = FLT_MAX, bd1 = FLT_MAX;
    int bi0 = -1, bi1 = -1;
    for (int c0 = 0; c0 < n_cols; c0 += EX_TROWS) {
        const int nr = min(EX_TROWS, n_cols - c0);
        __syncthreads();
        for (int e = threadIdx.x; e < nr * (DIM / 4); e += EX_THREADS) {
            const int r = e / (DIM / 4), k = e % (DIM / 4);
            tile[r][k] = reinterpret_cast<const float4*>(T + (size_t)(u.t_row0_f32 + c0 + r) * DIM)[k];
        }
        __syncthreads();
        for (int r = 0; r < nr; r++) {
            float acc[4][4];
            #pragma unroll
            for (int a = 0; a < 4; a++)
                #pragma unroll
                for (int l = 0; l < 4; l++) acc[a][l] = 0.f;
            #pragma unroll
            for (int j4 = 0; j4 < DIM / 4; j4++) {
                const float4 t = tile[r][j4];
                const int a = j4 & 3;
                float d;
                d = __fsub_rn(q[j4].x, t.x); acc[a][0] = __fadd_rn(acc[a][0], __fmul_rn(d, d));
                d = __fsub_rn(q[j4].y, t.y); acc[a][1] = __fadd_rn(acc[a][1], __fmul_rn(d, d));
                d = __fsub_rn(q[j4].z, t.z); acc[a][2] = __fadd_rn(acc[a][2], __fmul_rn(d, d));
                d = __fsub_rn(q[j4].w, t.w); acc[a][3] = __fadd_rn(acc[a][3], __fmul_rn(d, d));
            }
            float v[4];
            #pragma unroll
            for (int l = 0; l < 4; l++)
                v[l] = __fadd_rn(__fadd_rn(__fadd_rn(acc[0][l], acc[1][l]), acc[2][l]), acc[3][l]);
            const float s = __fadd_rn(__fadd_rn(v[0], v[2]), __fadd_rn(v[1], v[3]));
            const float dist = sqrtf(s);
            const int ti = u.t_local0 + c0 + r;
            if (dist < bd1) {                 // NaN / inf / >= FLT_MAX never enter (App. A.1)
                if (dist < bd0) { bd1 = bd0; bi1 = bi0; bd0 = dist; bi0 = ti; }
                else            { bd1 = dist; bi1 = ti; }
            }
        }
    }
    Top2 out;
    out.d1 = bi0 >= 0 ? bd0 : INFINITY; out.i1 = bi0;
    out.d2 = bi1 >= 0 ? bd1 : INFINITY; out.i2 = bi1;
    parts[(size_t)u.part_slot * TILE_M + threadIdx.x] = out;
}

void launch_match_exact(const float* Q, int n_query, const float* T, const MatchUnit* units, int n_units,
                        Top2* parts, const int* run_if_flag, cudaStream_t st)
{
    if (n_units <= 0) return;
    match_exact_kernel<<<n_units, EX_THREADS, 0, st>>>(Q, n_query, T, units, parts, run_if_flag);
}

// ---- merge: lexicographic (distance, train index), the order that reproduces OpenCV's tie rule ----
__device__ __forceinline__ bool lex_less(float da, int ia, float db, int ib)
{
    return da < db || (da == db && ia < ib);
}

__device__ __forceinline__ void top2_insert(float& d1, int& i1, float& d2, int& i2, float d, int i)
{
    if (i < 0) return;
    if (lex_less(d, i, d1, i1) || i1 < 0) {
        if (i1 >= 0) { d2 = d1; i2 = i1; }
        d1 = d; i1 = i;
    } else if (i2 < 0 || lex_less(d, i, d2, i2)) {
        d2 = d; i2 = i;
    }
}

__device__ __forceinline__ void top2_emit(float d1, int i1, float d2, int i2, float ratio, size_t row,
                                          int32_t* idx, float* dist, uint8_t* accept)
{
    idx[2 * row] = i1; idx[2 * row + 1] = i2;
    dist[2 * row] = i1 >= 0 ? d1 : 0.f; dist[2 * row + 1] = i2 >= 0 ? d2 : 0.f;
    if (accept) accept[row] = (i1 >= 0 && i2 >= 0 && d1 < __fmul_rn(ratio, d2)) ? 1 : 0;   // :67, App. A.4
}

// one thread per (segment, query row)
__global__ void merge_kernel(const Top2* __restrict__ parts, const MergeEntry* __restrict__ dir,
                             int n_segments, int n_rowblocks, int n_query, float ratio,
                             int32_t* __restrict__ idx, float* __restrict__ dist, uint8_t* __restrict__ accept)
{
    const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (gid >= (int64_t)n_segments * n_query) return;
    const int seg = (int)(gid / n_query), row = (int)(gid % n_query);
    const MergeEntry e = dir[seg * n_rowblocks + row / TILE_M];
    float d1 = INFINITY, d2 = INFINITY;
    int i1 = -1, i2 = -1;
    for (int s = 0; s < e.n_slots; s++) {
        const Top2 p = parts[(size_t)(e.first_slot + s) * TILE_M + (row % TILE_M)];
        top2_insert(d1, i1, d2, i2, p.d1, p.i1);
        top2_insert(d1, i1, d2, i2, p.d2, p.i2);
    }
    top2_emit(d1, i1, d2, i2, ratio, (size_t)gid, idx, dist, accept);
}

void launch_merge(const Top2* parts, const MergeEntry* dir, int n_segments, int n_rowblocks, int n_query,
                  float ratio, int32_t* idx, float* dist, uint8_t* accept, cudaStream_t st)
{
    const int64_t n = (int64_t)n_segments * n_query;
    if (n <= 0) return;
    merge_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(parts, dir, n_segments, n_rowblocks, n_query,
                                                              ratio, idx, dist, accept);
}

// merge of n_parts partial results laid out [n_parts][n_query][2] (multi-GPU train-tile shards)
__global__ void merge_parts_kernel(const float* __restrict__ dparts, const int32_t* __restrict__ iparts,
                                   int n_parts, int n_query, float ratio, int32_t* __restrict__ idx,
                                   float* __restrict__ dist, uint8_t* __restrict__ accept)
{
    const int row = blockIdx.x * blockDim.x + threadIdx.x;
    if (row >= n_query) return;
    float d1 = INFINITY, d2 = INFINITY;
    int i1 = -1, i2 = -1;
    for (int p = 0; p < n_parts; p++) {
        const size_t o = ((size_t)p * n_query + row) * 2;
        top2_insert(d1, i1, d2, i2, dparts[o], iparts[o]);
        top2_insert(d1, i1, d2, i2, dparts[o + 1], iparts[o + 1]);
    }
    top2_emit(d1, i1, d2, i2, ratio, (size_t)row, idx, dist, accept);
}

void launch_merge_parts(const float* dist_parts, const int32_t* idx_parts, int n_parts, int n_query, float ratio,
                        int32_t* idx, float* dist, uint8_t* accept, cudaStream_t st)
{
    if (n_query <= 0) return;
    merge_parts_kernel<<<(n_query + 255) / 256, 256, 0, st>>>(dist_parts, idx_parts, n_parts, n_query, ratio,
                                                              idx, dist, accept);
}

// final (idx,dist) arrays -> shard-exchange layout with global train indices
__global__ void shift_index_kernel(const int32_t* __restrict__ idx_in, const float* __restrict__ dist_in,
                                   int n_query, int32_t idx_base, float* __restrict__ dist, int32_t* __restrict__ idx)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n_query) return;
    const int v = idx_in[i];
    idx[i] = v >= 0 ? v + idx_base : -1;
    dist[i] = v >= 0 ? dist_in[i] : INFINITY;
}

void launch_shift_index(const int32_t* idx_in, const float* dist_in, int n_query, int32_t idx_base,
                        float* dist, int32_t* idx, cudaStream_t st)
{
    if (n_query <= 0) return;
    shift_index_kernel<<<(2 * n_query + 255) / 256, 256, 0, st>>>(idx_in, dist_in, n_query, idx_base, dist, idx);
}

}  // namespace cvg
