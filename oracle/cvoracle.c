/*
 * cvoracle.c — CPU restatement of OpenCV's BFMatcher(NORM_L2).knnMatch(k=2) and
 * findHomography(RANSAC) as called by the reference (src/TestsDetector.cpp:36,59-84).
 * TEST INFRASTRUCTURE ONLY (see cvoracle.h).  Pinned to cv2 4.13.0 through tests/golden/.
 *
 * OpenCV is an un-vendored dependency of the reference, so the functions below restate the
 * published algorithms of (paths inside the OpenCV source tree, not on disk here):
 *   modules/features2d/src/matchers.cpp   BFMatcher::knnMatchImpl
 *   modules/core/src/batch_distance.cpp   batchDistance / normL2Sqr_ (fp32, 4 lanes x 4 accumulators)
 *   modules/calib3d/src/ptsetreg.cpp      RANSACPointSetRegistrator, RANSACUpdateNumIters
 *   modules/calib3d/src/fundam.cpp        HomographyEstimatorCallback, HomographyRefineCallback,
 *                                         findHomography
 *   modules/calib3d/src/levmarq.cpp       LMSolverImpl::run
 *   modules/core/src/lapack.cpp           JacobiImpl_, hypot, solve/invert(DECOMP_EIG), SVBkSb
 *   modules/core/include/.../core.hpp     RNG (multiply-with-carry)
 * as summarised in SURVEY.md App. A, B, D.  Build with -ffp-contract=off (no FMA anywhere).
 */
#include "cvoracle.h"
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------ */
/* Match stage — SURVEY App. A                                                                */
/* ------------------------------------------------------------------------------------------ */

/* App. A.3: 16 partial sums acc[a][l], element j -> a=(j/4)%4, l=j%4, separate mul and add. */
float cvo_l2sqr(const float* a, const float* b, int dim)
{
    float acc[4][4];
    memset(acc, 0, sizeof acc);
    int j = 0;
    for (; j + 16 <= dim; j += 16)
        for (int u = 0; u < 4; u++)
            for (int l = 0; l < 4; l++) {
                float d = a[j + 4 * u + l] - b[j + 4 * u + l];
                float p = d * d;
                acc[u][l] = acc[u][l] + p;
            }
    float v[4];
    for (int l = 0; l < 4; l++) v[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
    /* tail (dim % 16): 4-lane blocks into accumulator 0's lanes, then scalars; D=128 never gets here */
    for (; j + 4 <= dim; j += 4)
        for (int l = 0; l < 4; l++) {
            float d = a[j + l] - b[j + l];
            v[l] = v[l] + d * d;
        }
    float s = (v[0] + v[2]) + (v[1] + v[3]);
    for (; j < dim; j++) {
        float d = a[j] - b[j];
        s = s + d * d;
    }
    return s;
}

static inline int32_t f2i(float f) { int32_t i; memcpy(&i, &f, 4); return i; }

/* App. A.1/A.2: K-list initialised to FLT_MAX, int-bit-pattern compare, strict <, stable. */
static void knn2_row(const float* q, const float* t, int nt, int dim, int32_t* idx, float* dist)
{
    float bd[2] = { FLT_MAX, FLT_MAX };
    int32_t bi[2] = { -1, -1 };
    for (int j = 0; j < nt; j++) {
        float d = sqrtf(cvo_l2sqr(q, t + (size_t)j * dim, dim));
        int32_t di = f2i(d);
        if (di < f2i(bd[1])) {
            if (di < f2i(bd[0])) {
                bd[1] = bd[0]; bi[1] = bi[0];
                bd[0] = d;     bi[0] = j;
            } else {
                bd[1] = d; bi[1] = j;
            }
        }
    }
    idx[0] = bi[0]; idx[1] = bi[1];
    dist[0] = bd[0]; dist[1] = bd[1];
}

typedef struct { const float* q; const float* t; int nq, nt, dim; int32_t* idx; float* dist; int tid, nth; } knn_job;

static void* knn_worker(void* arg)
{
    knn_job* j = (knn_job*)arg;
    /* interleaved blocks of 16 query rows per thread; results do not depend on the split (App. A.5) */
    for (int b = j->tid * 16; b < j->nq; b += j->nth * 16)
        for (int i = b; i < b + 16 && i < j->nq; i++)
            knn2_row(j->q + (size_t)i * j->dim, j->t, j->nt, j->dim, j->idx + 2 * i, j->dist + 2 * i);
    return NULL;
}

void cvo_knn2(const float* q, int nq, const float* t, int nt, int dim,
              int32_t* idx, float* dist, int nthreads)
{
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 256) nthreads = 256;
    pthread_t th[256];
    knn_job jobs[256];
    for (int k = 0; k < nthreads; k++) {
        knn_job jb = { q, t, nq, nt, dim, idx, dist, k, nthreads };
        jobs[k] = jb;
        if (k > 0) pthread_create(&th[k], NULL, knn_worker, &jobs[k]);
    }
    knn_worker(&jobs[0]);
    for (int k = 1; k < nthreads; k++) pthread_join(th[k], NULL);
}

void cvo_ratio(const int32_t* idx, const float* dist, int nq, float ratio, uint8_t* accept)
{
    for (int i = 0; i < nq; i++) {
        int two = idx[2 * i] >= 0 && idx[2 * i + 1] >= 0;
        float thr = ratio * dist[2 * i + 1];
        accept[i] = (uint8_t)(two && dist[2 * i] < thr);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Jacobi eigen-solver — SURVEY App. D.1 / D.2                                                */
/* ------------------------------------------------------------------------------------------ */

static double cv_hypot(double a, double b)
{
    a = fabs(a); b = fabs(b);
    if (a > b) { b /= a; return a * sqrt(1 + b * b); }
    if (b > 0) { a /= b; return b * sqrt(1 + a * a); }
    return 0;
}

#define JMAXN 16

int cvo_jacobi(double* A, double* W, double* V, int n)
{
    const double eps = DBL_EPSILON;
    int indR[JMAXN], indC[JMAXN];
    int i, j, k, m;
    double mv;
    if (n > JMAXN) return 0;

    for (i = 0; i < n; i++) {
        for (j = 0; j < n; j++) V[i * n + j] = 0;
        V[i * n + i] = 1;
    }
    for (k = 0; k < n; k++) {
        W[k] = A[(n + 1) * k];
        if (k < n - 1) {
            for (m = k + 1, mv = fabs(A[n * k + m]), i = k + 2; i < n; i++) {
                double val = fabs(A[n * k + i]);
                if (mv < val) mv = val, m = i;
            }
            indR[k] = m;
        }
        if (k > 0) {
            for (m = 0, mv = fabs(A[k]), i = 1; i < k; i++) {
                double val = fabs(A[n * i + k]);
                if (mv < val) mv = val, m = i;
            }
            indC[k] = m;
        }
    }

    int iters, maxIters = n * n * 30;
    if (n > 1) for (iters = 0; iters < maxIters; iters++) {
        for (k = 0, mv = fabs(A[indR[0]]), i = 1; i < n - 1; i++) {
            double val = fabs(A[n * i + indR[i]]);
            if (mv < val) mv = val, k = i;
        }
        int l = indR[k];
        for (i = 1; i < n; i++) {
            double val = fabs(A[n * indC[i] + i]);
            if (mv < val) mv = val, k = indC[i], l = i;
        }
        double p = A[n * k + l];
        if (fabs(p) <= eps) break;
        double y = (W[l] - W[k]) * 0.5;
        double t = fabs(y) + cv_hypot(p, y);
        double s = cv_hypot(p, t);
        double c = t / s;
        s = p / s; t = (p / t) * p;
        if (y < 0) s = -s, t = -t;
        A[n * k + l] = 0;
        W[k] -= t;
        W[l] += t;

        double a0, b0;
#define ROT(v0, v1) a0 = v0, b0 = v1, v0 = a0 * c - b0 * s, v1 = a0 * s + b0 * c
        for (i = 0; i < k; i++)     ROT(A[n * i + k], A[n * i + l]);
        for (i = k + 1; i < l; i++) ROT(A[n * k + i], A[n * i + l]);
        for (i = l + 1; i < n; i++) ROT(A[n * k + i], A[n * l + i]);
        for (i = 0; i < n; i++)     ROT(V[n * k + i], V[n * l + i]);
#undef ROT
        for (j = 0; j < 2; j++) {
            int idx = j == 0 ? k : l;
            if (idx < n - 1) {
                for (m = idx + 1, mv = fabs(A[n * idx + m]), i = idx + 2; i < n; i++) {
                    double val = fabs(A[n * idx + i]);
                    if (mv < val) mv = val, m = i;
                }
                indR[idx] = m;
            }
            if (idx > 0) {
                for (m = 0, mv = fabs(A[idx]), i = 1; i < idx; i++) {
                    double val = fabs(A[n * i + idx]);
                    if (mv < val) mv = val, m = i;
                }
                indC[idx] = m;
            }
        }
    }
    /* sort eigenvalues (descending) & eigenvectors */
    for (k = 0; k < n - 1; k++) {
        m = k;
        for (i = k + 1; i < n; i++)
            if (W[m] < W[i]) m = i;
        if (k != m) {
            double tmp = W[m]; W[m] = W[k]; W[k] = tmp;
            for (i = 0; i < n; i++) {
                tmp = V[n * m + i]; V[n * m + i] = V[n * k + i]; V[n * k + i] = tmp;
            }
        }
    }
    return 1;
}

/* cv::solve(A, b, x, DECOMP_EIG) for square symmetric A, single right-hand side:
 * Jacobi, then SVBkSb with threshold = 2*DBL_EPSILON*sum(w). */
int cvo_solve_eig(const double* A, const double* b, double* x, int n)
{
    double a[JMAXN * JMAXN], w[JMAXN], v[JMAXN * JMAXN];
    if (n > JMAXN) return 0;
    memcpy(a, A, sizeof(double) * n * n);
    cvo_jacobi(a, w, v, n);
    double threshold = 0;
    int i, j;
    for (i = 0; i < n; i++) x[i] = 0;
    for (i = 0; i < n; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (i = 0; i < n; i++) {
        double wi = w[i];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        double s = 0;
        for (j = 0; j < n; j++) s += v[i * n + j] * b[j];
        s *= wi;
        for (j = 0; j < n; j++) x[j] = x[j] + s * v[i * n + j];
    }
    return 1;
}

/* cv::invert(A, Ainv, DECOMP_EIG): eigen + SVD::backSubst with identity rhs. */
double cvo_invert_eig(const double* A, double* Ainv, int n)
{
    double a[JMAXN * JMAXN], w[JMAXN], v[JMAXN * JMAXN], buf[JMAXN];
    if (n > JMAXN) return 0;
    memcpy(a, A, sizeof(double) * n * n);
    cvo_jacobi(a, w, v, n);
    double threshold = 0;
    int i, j, k;
    for (i = 0; i < n * n; i++) Ainv[i] = 0;
    for (i = 0; i < n; i++) threshold += w[i];
    threshold *= DBL_EPSILON * 2;
    for (k = 0; k < n; k++) {
        double wi = w[k];
        if (fabs(wi) <= threshold) continue;
        wi = 1 / wi;
        for (j = 0; j < n; j++) buf[j] = v[k * n + j] * wi;
        for (i = 0; i < n; i++)
            for (j = 0; j < n; j++)
                Ainv[i * n + j] = Ainv[i * n + j] + v[k * n + i] * buf[j];
    }
    return w[0] >= DBL_EPSILON ? w[n - 1] / w[0] : 0;
}

/* ------------------------------------------------------------------------------------------ */
/* HomographyEstimatorCallback — SURVEY App. B.4 / B.5 / B.6                                  */
/* ------------------------------------------------------------------------------------------ */

static void mat3mul(const double* a, const double* b, double* d)
{
    for (int i = 0; i < 3; i++)
        for (int j = 0; j < 3; j++)
            d[i * 3 + j] = (a[i * 3 + 0] * b[0 * 3 + j] + a[i * 3 + 1] * b[1 * 3 + j]) + a[i * 3 + 2] * b[2 * 3 + j];
}

int cvo_run_kernel(const float* M, const float* m, int count, double Hout[9])
{
    double LtL[81], W[9], V[81];
    double cMx = 0, cMy = 0, cmx = 0, cmy = 0, sMx = 0, sMy = 0, smx = 0, smy = 0;
    int i, j, k;
    for (i = 0; i < count; i++) {
        cmx += m[2 * i]; cmy += m[2 * i + 1];
        cMx += M[2 * i]; cMy += M[2 * i + 1];
    }
    cmx /= count; cmy /= count; cMx /= count; cMy /= count;
    for (i = 0; i < count; i++) {
        smx += fabs(m[2 * i] - cmx);
        smy += fabs(m[2 * i + 1] - cmy);
        sMx += fabs(M[2 * i] - cMx);
        sMy += fabs(M[2 * i + 1] - cMy);
    }
    if (fabs(smx) < DBL_EPSILON || fabs(smy) < DBL_EPSILON ||
        fabs(sMx) < DBL_EPSILON || fabs(sMy) < DBL_EPSILON)
        return 0;
    smx = count / smx; smy = count / smy;
    sMx = count / sMx; sMy = count / sMy;

    double invHnorm[9] = { 1. / smx, 0, cmx, 0, 1. / smy, cmy, 0, 0, 1 };
    double Hnorm2[9] = { sMx, 0, -cMx * sMx, 0, sMy, -cMy * sMy, 0, 0, 1 };

    memset(LtL, 0, sizeof LtL);
    for (i = 0; i < count; i++) {
        double x = (m[2 * i] - cmx) * smx, y = (m[2 * i + 1] - cmy) * smy;
        double X = (M[2 * i] - cMx) * sMx, Y = (M[2 * i + 1] - cMy) * sMy;
        double Lx[9] = { X, Y, 1, 0, 0, 0, -x * X, -x * Y, -x };
        double Ly[9] = { 0, 0, 0, X, Y, 1, -y * X, -y * Y, -y };
        for (j = 0; j < 9; j++)
            for (k = j; k < 9; k++)
                LtL[j * 9 + k] += Lx[j] * Lx[k] + Ly[j] * Ly[k];
    }
    for (j = 0; j < 9; j++)
        for (k = 0; k < j; k++)
            LtL[j * 9 + k] = LtL[k * 9 + j];

    cvo_jacobi(LtL, W, V, 9);
    double Htemp[9], H0[9];
    mat3mul(invHnorm, V + 72, Htemp);
    mat3mul(Htemp, Hnorm2, H0);
    double sc = 1. / H0[8];
    for (i = 0; i < 9; i++) Hout[i] = H0[i] * sc;
    return 1;
}

static int have_collinear(const float* p, int count)
{
    int j, k, i = count - 1;
    for (j = 0; j < i; j++) {
        /* Point2f members are float: the subtraction is fp32, then widened. */
        double dx1 = (float)(p[2 * j] - p[2 * i]);
        double dy1 = (float)(p[2 * j + 1] - p[2 * i + 1]);
        for (k = 0; k < j; k++) {
            double dx2 = (float)(p[2 * k] - p[2 * i]);
            double dy2 = (float)(p[2 * k + 1] - p[2 * i + 1]);
            if (fabs(dx2 * dy1 - dy2 * dx1) <= FLT_EPSILON * (fabs(dx1) + fabs(dy1) + fabs(dx2) + fabs(dy2)))
                return 1;
        }
    }
    return 0;
}

static double det3_pts(const float* p, int a, int b, int c)
{
    double a00 = p[2 * a], a01 = p[2 * a + 1], a02 = 1.;
    double a10 = p[2 * b], a11 = p[2 * b + 1], a12 = 1.;
    double a20 = p[2 * c], a21 = p[2 * c + 1], a22 = 1.;
    return a00 * (a11 * a22 - a21 * a12) - a01 * (a10 * a22 - a20 * a12) + a02 * (a10 * a21 - a20 * a11);
}

int cvo_check_subset(const float* ms1, const float* ms2, int count)
{
    if (have_collinear(ms1, count) || have_collinear(ms2, count)) return 0;
    if (count == 4) {
        static const int tt[4][3] = { {0, 1, 2}, {1, 2, 3}, {0, 2, 3}, {0, 1, 3} };
        int negative = 0;
        for (int i = 0; i < 4; i++) {
            double dA = det3_pts(ms1, tt[i][0], tt[i][1], tt[i][2]);
            double dB = det3_pts(ms2, tt[i][0], tt[i][1], tt[i][2]);
            negative += dA * dB < 0;
        }
        if (negative != 0 && negative != 4) return 0;
    }
    return 1;
}

void cvo_compute_error(const float* M, const float* m, int n, const double H[9], float* err)
{
    float Hf[8];
    for (int i = 0; i < 8; i++) Hf[i] = (float)H[i];
    for (int i = 0; i < n; i++) {
        float Mx = M[2 * i], My = M[2 * i + 1];
        float ww = 1.f / (Hf[6] * Mx + Hf[7] * My + 1.f);
        float dx = (Hf[0] * Mx + Hf[1] * My + Hf[2]) * ww - m[2 * i];
        float dy = (Hf[3] * Mx + Hf[4] * My + Hf[5]) * ww - m[2 * i + 1];
        err[i] = dx * dx + dy * dy;
    }
}

int cvo_update_num_iters(double p, double ep, int model_points, int max_iters)
{
    p = p > 0. ? p : 0.;   p = p < 1. ? p : 1.;
    ep = ep > 0. ? ep : 0.; ep = ep < 1. ? ep : 1.;
    double num = 1. - p > DBL_MIN ? 1. - p : DBL_MIN;
    double denom = 1. - pow(1. - ep, model_points);
    if (denom < DBL_MIN) return 0;
    num = log(num);
    denom = log(denom);
    return denom >= 0 || -num >= max_iters * (-denom) ? max_iters : (int)lrint(num / denom);
}

uint32_t cvo_rng_next(uint64_t* state)
{
    *state = (uint64_t)(uint32_t)(*state) * 4164903690U + (uint32_t)(*state >> 32);
    return (uint32_t)(*state);
}

/* App. D.3: draw 4 distinct indices (redraw on duplicates), accept iff checkSubset. */
static int get_subset(const float* src, const float* dst, int n, uint64_t* rng, int max_attempts,
                      int idx[4], float ms1[8], float ms2[8], int64_t* draws)
{
    for (int it = 0; it < max_attempts; it++) {
        int i;
        for (i = 0; i < 4; i++) {
            int v, dup;
            do {
                v = (int)(cvo_rng_next(rng) % (uint32_t)n);
                (*draws)++;
                dup = 0;
                for (int k = 0; k < i; k++) dup |= idx[k] == v;
            } while (dup);
            idx[i] = v;
            ms1[2 * i] = src[2 * v]; ms1[2 * i + 1] = src[2 * v + 1];
            ms2[2 * i] = dst[2 * v]; ms2[2 * i + 1] = dst[2 * v + 1];
        }
        if (cvo_check_subset(ms1, ms2, 4)) return 1;
    }
    return 0;
}

int cvo_ransac_stage(const float* src, const float* dst, int n, double thr, double conf,
                     int max_iters, double Hbest[9], uint8_t* best_mask,
                     int32_t* samples, int32_t* counts, cvo_ransac_info* info)
{
    cvo_ransac_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof *info);
    info->best_iter = -1;
    int niters = max_iters > 1 ? max_iters : 1;
    if (samples) for (int i = 0; i < niters * 4; i++) samples[i] = -1;
    if (counts)  for (int i = 0; i < niters; i++) counts[i] = -1;
    info->niters_final = niters;
    if (n < 4) return 0;
    if (n == 4) {
        if (!cvo_run_kernel(src, dst, 4, Hbest)) return 0;
        memset(best_mask, 1, 4);
        info->best_count = 4;
        return 1;
    }
    uint64_t rng = (uint64_t)-1;
    float t = (float)(thr * thr);
    float* err = (float*)malloc(sizeof(float) * n);
    uint8_t* mask = (uint8_t*)malloc(n);
    int max_good = 0, iter;
    double H[9];
    for (iter = 0; iter < niters; iter++) {
        int idx[4]; float ms1[8], ms2[8];
        if (!get_subset(src, dst, n, &rng, 10000, idx, ms1, ms2, &info->draws)) {
            if (iter == 0) { free(err); free(mask); info->iters_run = 0; return 0; }
            break;
        }
        if (samples) for (int k = 0; k < 4; k++) samples[4 * iter + k] = idx[k];
        if (!cvo_run_kernel(ms1, ms2, 4, H)) continue;
        cvo_compute_error(src, dst, n, H, err);
        int good = 0;
        for (int i = 0; i < n; i++) { mask[i] = (uint8_t)(err[i] <= t); good += mask[i]; }
        if (counts) counts[iter] = good;
        if (good > (max_good > 3 ? max_good : 3)) {
            memcpy(best_mask, mask, n);
            memcpy(Hbest, H, sizeof H);
            max_good = good;
            info->best_iter = iter;
            niters = cvo_update_num_iters(conf, (double)(n - good) / n, 4, niters);
        }
    }
    info->iters_run = iter;
    info->niters_final = niters;
    info->best_count = max_good;
    free(err); free(mask);
    if (max_good > 0) return 1;
    memset(best_mask, 0, n);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* LM refinement — SURVEY App. B.8 (cv::LMSolver with HomographyRefineCallback, 10 iterations)*/
/* ------------------------------------------------------------------------------------------ */


typedef struct { const float* M; const float* m; int n; } lm_pts;
typedef void (*lm_callback)(const lm_pts* p, const double* h, double* r, double* J);

/* HomographyRefineCallback::compute as built in cv2 4.13.0: NINE parameters (the binary carries the
 * assertion "J.isContinuous() && J.cols == 9"), i.e. h[8] is optimised too and the scale gauge is
 * left to solve(DECOMP_EIG)'s eigenvalue truncation; H is renormalised by h[8] afterwards.
 * residuals r[2n], optional Jacobian J[2n*9]. */
static void homography_refine_compute(const lm_pts* p, const double* h, double* r, double* J)
{
    const float *M = p->M, *m = p->m;
    for (int i = 0; i < p->n; i++) {
        double Mx = M[2 * i], My = M[2 * i + 1];
        double ww = h[6] * Mx + h[7] * My + h[8];
        ww = fabs(ww) > DBL_EPSILON ? 1. / ww : 0;
        double xi = (h[0] * Mx + h[1] * My + h[2]) * ww;
        double yi = (h[3] * Mx + h[4] * My + h[5]) * ww;
        r[2 * i] = xi - m[2 * i];
        r[2 * i + 1] = yi - m[2 * i + 1];
        if (J) {
            double* Jp = J + 18 * i;
            Jp[0] = Mx * ww; Jp[1] = My * ww; Jp[2] = ww;
            Jp[3] = Jp[4] = Jp[5] = 0.;
            Jp[6] = -Mx * ww * xi; Jp[7] = -My * ww * xi; Jp[8] = -ww * xi;
            Jp[9] = Jp[10] = Jp[11] = 0.;
            Jp[12] = Mx * ww; Jp[13] = My * ww; Jp[14] = ww;
            Jp[15] = -Mx * ww * yi; Jp[16] = -My * ww * yi; Jp[17] = -ww * yi;
        }
    }
}

/* Affine2DRefineCallback::compute (cv::estimateAffine2D's refine step).  Not on the reference's
 * path: it is here because estimateAffine2D exposes refineIters, which lets the tests observe
 * cv2's LMSolver iterate by iterate and pin lm_run() bit-for-bit. */
static void affine_refine_compute(const lm_pts* p, const double* h, double* r, double* J)
{
    const float *M = p->M, *m = p->m;
    for (int i = 0; i < p->n; i++) {
        double Mx = M[2 * i], My = M[2 * i + 1];
        double xi = h[0] * Mx + h[1] * My + h[2];
        double yi = h[3] * Mx + h[4] * My + h[5];
        r[2 * i] = xi - m[2 * i];
        r[2 * i + 1] = yi - m[2 * i + 1];
        if (J) {
            double* Jp = J + 12 * i;
            Jp[0] = Mx; Jp[1] = My; Jp[2] = 1.;
            Jp[3] = Jp[4] = Jp[5] = 0.;
            Jp[6] = Jp[7] = Jp[8] = 0.;
            Jp[9] = Mx; Jp[10] = My; Jp[11] = 1.;
        }
    }
}
/* cv::norm(r, NORM_L2SQR) for CV_64F as built in cv2 4.13.0 (AVX2 dispatch; pinned by probing
 * cv2.norm on this image, 414/414 random vectors): blocks of 16 into 4 lanes x 4 accumulators with
 * FMA, r = ((a0+a1)+a2)+a3, s = (r0+r1)+(r2+r3); then groups of 4 with separate multiply and
 * in-order adds; then a <4 scalar tail with FMA.  The LM accept test "Sd < S" is decided by rounding
 * noise near convergence, so this order is load-bearing for H at the 1e-8 level. */
static double norm_l2sqr(const double* a, int n)
{
    double acc[4][4];
    memset(acc, 0, sizeof acc);
    int j = 0;
    for (; j <= n - 16; j += 16)
        for (int k = 0; k < 4; k++)
            for (int l = 0; l < 4; l++) {
                double v = a[j + 4 * k + l];
                acc[k][l] = fma(v, v, acc[k][l]);
            }
    double r[4];
    for (int l = 0; l < 4; l++) r[l] = ((acc[0][l] + acc[1][l]) + acc[2][l]) + acc[3][l];
    double s = (r[0] + r[1]) + (r[2] + r[3]);
    s = 0.0 + s;
    for (; j <= n - 4; j += 4)
        for (int q = 0; q < 4; q++) s = s + a[j + q] * a[j + q];
    for (; j < n; j++) s = fma(a[j], a[j], s);
    return s;
}
/* cv::gemm inner product for small matrices (GEMMSingleMul, A*Bt form): 4 interleaved accumulators,
 * no FMA, ((s0+s1)+s2)+s3 (pinned by probing cv2.gemm, rows < 100).  From 100 rows on OpenCV hands
 * J^T r to OpenBLAS (HAL threshold), whose order is machine-specific: sequential is used there and
 * is NOT pinned. */
static double dot4acc(const double* a, int sa, const double* b, int sb, int n)
{
    double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
    int k = 0;
    for (; k <= n - 4; k += 4) {
        s0 += a[(k) * sa] * b[(k) * sb];
        s1 += a[(k + 1) * sa] * b[(k + 1) * sb];
        s2 += a[(k + 2) * sa] * b[(k + 2) * sb];
        s3 += a[(k + 3) * sa] * b[(k + 3) * sb];
    }
    for (; k < n; k++) s0 += a[k * sa] * b[k * sb];
    return ((s0 + s1) + s2) + s3;
}
/* cv::Mat::dot generic path (dotProd_): unrolled by 4, no FMA */
static double dot_unrolled(const double* a, const double* b, int n)
{
    double r = 0;
    int i = 0;
    for (; i <= n - 4; i += 4)
        r += ((a[i] * b[i] + a[i + 1] * b[i + 1]) + a[i + 2] * b[i + 2]) + a[i + 3] * b[i + 3];
    for (; i < n; i++) r += a[i] * b[i];
    return r;
}
static double norm_inf(const double* r, int n)
{
    double s = 0;
    for (int i = 0; i < n; i++) { double a = fabs(r[i]); if (a > s) s = a; }
    return s;
}
/* A = J^T J (lx x lx, cv::mulTransposed: sequential over rows, no FMA), v = J^T r (cv::gemm) */
static void normal_eqs(const double* J, const double* r, int rows, int lx, double* A, double* v)
{
    for (int i = 0; i < lx; i++) {
        for (int j = i; j < lx; j++) {
            double s = 0;
            for (int k = 0; k < rows; k++) s += J[k * lx + i] * J[k * lx + j];
            A[i * lx + j] = s;
        }
        if (rows < 100) {
            v[i] = dot4acc(J + i, lx, r, 1, rows);
        } else {
            double s = 0;
            for (int k = 0; k < rows; k++) s += J[k * lx + i] * r[k];
            v[i] = s;
        }
    }
    for (int i = 0; i < lx; i++)
        for (int j = 0; j < i; j++) A[i * lx + j] = A[j * lx + i];
}

/* cv::LMSolverImpl::run (calib3d/src/levmarq.cpp; Balda's LMFsolve schedule), eps = FLT_EPSILON */
static int lm_run(lm_callback cb, const lm_pts* pts, double* x, int lx, int max_iters)
{
    const int rows = 2 * pts->n;
    const double epsx = FLT_EPSILON, epsf = FLT_EPSILON;
    double xd[9], A[81], Ap[81], v[9], d[9], D[9], temp_d[9];
    double* r = (double*)malloc(sizeof(double) * rows);
    double* rd = (double*)malloc(sizeof(double) * rows);
    double* J = (double*)malloc(sizeof(double) * rows * lx);
    int i, iter = 0;

    cb(pts, x, r, J);
    double S = norm_l2sqr(r, rows);
    normal_eqs(J, r, rows, lx, A, v);
    for (i = 0; i < lx; i++) D[i] = A[i * lx + i];

    const double Rlo = 0.25, Rhi = 0.75;
    double lambda = 1, lc = 0.75;

    for (;;) {
        memcpy(Ap, A, sizeof(double) * lx * lx);
        for (i = 0; i < lx; i++) Ap[i * lx + i] += lambda * D[i];
        cvo_solve_eig(Ap, v, d, lx);
        for (i = 0; i < lx; i++) xd[i] = x[i] - d[i];
        cb(pts, xd, rd, NULL);
        double Sd = norm_l2sqr(rd, rows);
        /* temp_d = -A*d + 2*v ; dS = d . temp_d */
        for (i = 0; i < lx; i++) {
            double s = dot4acc(A + i * lx, 1, d, 1, lx);
            temp_d[i] = -1 * s + 2 * v[i];
        }
        double dS = dot_unrolled(d, temp_d, lx);
        double R = (S - Sd) / (fabs(dS) > DBL_EPSILON ? dS : 1);

        if (R > Rhi) {
            lambda *= 0.5;
            if (lambda < lc) lambda = 0;
        } else if (R < Rlo) {
            double t = dot_unrolled(d, v, lx);
            double nu = (Sd - S) / (fabs(t) > DBL_EPSILON ? t : 1) + 2;
            nu = nu > 2. ? nu : 2.;
            nu = nu < 10. ? nu : 10.;
            if (lambda == 0) {
                cvo_invert_eig(A, Ap, lx);
                double maxval = DBL_EPSILON;
                for (i = 0; i < lx; i++) {
                    double a = fabs(Ap[i * lx + i]);
                    if (a > maxval) maxval = a;
                }
                lambda = lc = 1. / maxval;
                nu *= 0.5;
            }
            lambda *= nu;
        }

        if (Sd < S) {
            S = Sd;
            memcpy(x, xd, sizeof(double) * lx);
            cb(pts, x, r, J);
            normal_eqs(J, r, rows, lx, A, v);
        }
        iter++;
        int proceed = iter < max_iters && norm_inf(d, lx) >= epsx && norm_inf(r, rows) >= epsf;
        if (!proceed) break;
    }
    free(r); free(rd); free(J);
    return iter == max_iters ? -iter : iter;
}

int cvo_lm_refine(const float* M, const float* m, int n, double H[9], int max_iters)
{
    lm_pts p = { M, m, n };
    int it = lm_run(homography_refine_compute, &p, H, 9, max_iters);
    double sc = 1. / H[8];
    for (int i = 0; i < 9; i++) H[i] = H[i] * sc;
    return it;
}

/* LM refine step of cv::estimateAffine2D on a 2x3 matrix (test handle, see affine_refine_compute) */
int cvo_lm_refine_affine(const float* M, const float* m, int n, double A6[6], int max_iters)
{
    lm_pts p = { M, m, n };
    return lm_run(affine_refine_compute, &p, A6, 6, max_iters);
}

/* ------------------------------------------------------------------------------------------ */
/* cv::findHomography — SURVEY App. B.1, B.7-B.9                                              */
/* ------------------------------------------------------------------------------------------ */

int cvo_find_homography(const float* src, const float* dst, int n, int method, double thr,
                        int max_iters, double conf, double H[9], uint8_t* mask, int* found,
                        uint8_t* ransac_mask, cvo_ransac_info* info)
{
    cvo_ransac_info local;
    if (!info) info = &local;
    memset(info, 0, sizeof *info);
    info->best_iter = -1;
    *found = 0;
    if (n < 4) return -1;
    if (thr <= 0) thr = 3;
    int result;
    uint8_t* tmask = (uint8_t*)malloc(n);
    if (method == 0 || n == 4) {
        memset(tmask, 1, n);
        result = cvo_run_kernel(src, dst, n, H) > 0;
    } else {
        result = cvo_ransac_stage(src, dst, n, thr, conf, max_iters, H, tmask, NULL, NULL, info);
    }
    if (ransac_mask) { if (result) memcpy(ransac_mask, tmask, n); else memset(ransac_mask, 0, n); }

    if (result && n > 4) {
        float* s1 = (float*)malloc(sizeof(float) * 2 * n);
        float* d1 = (float*)malloc(sizeof(float) * 2 * n);
        int np = 0;
        for (int i = 0; i < n; i++)
            if (tmask[i]) {
                s1[2 * np] = src[2 * i]; s1[2 * np + 1] = src[2 * i + 1];
                d1[2 * np] = dst[2 * i]; d1[2 * np + 1] = dst[2 * i + 1];
                np++;
            }
        info->n_refit = np;
        if (np > 0) {
            if (method != 0) cvo_run_kernel(s1, d1, np, H);
            info->lm_iters = cvo_lm_refine(s1, d1, np, H, 10);
            /* B.9 (cv2 4.13.0): the returned mask is recomputed from the refined H (also for method 0) */
            {
                float* err = (float*)malloc(sizeof(float) * n);
                float t = (float)(thr * thr);
                cvo_compute_error(src, dst, n, H, err);
                for (int i = 0; i < n; i++) tmask[i] = (uint8_t)(err[i] <= t);
                free(err);
            }
        }
        free(s1); free(d1);
    }
    if (result) {
        memcpy(mask, tmask, n);
        *found = 1;
    } else {
        memset(mask, 0, n);
        for (int i = 0; i < 9; i++) H[i] = 0;
    }
    free(tmask);
    return 0;
}

double cvo_det3(const double m[9])
{
    return m[0] * (m[4] * m[8] - m[5] * m[7]) - m[1] * (m[3] * m[8] - m[5] * m[6]) + m[2] * (m[3] * m[7] - m[4] * m[6]);
}

/* reference src/TestsDetector.cpp:74,79,81,84 */
int cvo_gate(int n_good, int found, const uint8_t* mask, const double H[9],
             int min_inliers, float det_lo, float det_hi)
{
    if (n_good < min_inliers) return CVO_LT4_MATCHES;
    if (!found) return CVO_H_EMPTY;
    int cnt = 0;
    for (int i = 0; i < n_good; i++) cnt += mask[i] != 0;
    if (cnt < min_inliers) return CVO_LT4_INLIERS;
    double detH = fabs(cvo_det3(H));
    if (detH < det_lo || detH > det_hi) return CVO_DET_REJECT;
    return CVO_ACCEPT;
}
