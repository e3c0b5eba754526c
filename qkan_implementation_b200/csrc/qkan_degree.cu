// qkan_degree.cu - SURVEY 8(f) rank 4: the degree-evaluation least squares of the reference's
// DegreeOptimizer.evaluate_degree (original_degree_optimizer/DegreeOptimizer.py:122-158) on the GPU.
//
// The reference builds, for every candidate degree d = 0..D, the design matrix
//     X_d = [T_0(x) | T_1(x) | ... | T_d(x)],   T_k(x) = cos(k arccos(clip(x, -1, 1)))    (ChebyshevStep.py:32-53)
// over all n samples and F features, solves min |X_d c - y| with np.linalg.lstsq and scores the fit.  All
// X_d are column prefixes of X_D, so ONE Gram matrix of the augmented matrix [X_D | y] carries the normal
// equations of every degree (leading blocks), X_d^T y, sum(y) and y^T y.  Two kernels:
//   * qkan_cheb_gram_kernel: fused Chebyshev-feature generation + FP64 tensor-core SYRK (mma.sync m8n8k4,
//     DMMA): a CTA owns one 64 x 64 tile of the upper triangle and one slice of the samples; features are
//     produced into shared memory from x (recurrence T_{k+1} = 2 x T_k - T_{k-1}) while the next chunk's x
//     is already in flight; partial tiles are reduced in a fixed order (deterministic, no atomics).
//     Internally columns are feature-major (f (D+1) + k) so that a 64-column tile reads only ~64 / (D+1)
//     features of x; the result is written degree-major (k F + f), the reference's np.hstack order.
//     For D >= 1 the F identical all-ones T_0 columns are kept once (GramParams); the reduce kernel sums the slices
//     into the matrix of the distinct columns and the expand kernel writes the full matrix of the C ABI from it.
//   * qkan_cheb_residual_tile_kernel (D <= 4, F <= 128) / qkan_cheb_residual_kernel (any shape, one warp per
//     sample): the fits' explicit residuals r_d = y - X_d c_d for all d in one pass, giving the sums behind the
//     reference's MSE / R^2 (DegreeOptimizer.py:277-312) without the cancellation of y^T y - 2 c^T g + c^T G c,
//     and X_D^T r_d for one step of iterative refinement.
// The (D+1) small dense solves run on the host above the C ABI (numpy), as in the reference.
#include "../../include/qkan_b200.h"

#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>

namespace {

constexpr int TILE = 64;          // Gram tile edge
constexpr int KC = 32;            // samples per shared-memory chunk
constexpr int LDK = KC + 4;       // shared-memory chunks are [column][sample] with 36 doubles per column: the 16 lanes of a
                                  // half warp read fragments (m, k) = (0..3, 0..3) -> words 36 m + k, 16 different banks, and
                                  // the producers write one column for 32 consecutive samples -> consecutive banks
constexpr int GRAM_THREADS = 128; // 4 warps, each a 32 x 32 quadrant of the tile
constexpr int MAX_D = 16;

__device__ __forceinline__ double clip_unit(double x) {
    return x < -1.0 ? -1.0 : (x > 1.0 ? 1.0 : x);       // comparisons keep NaN, like np.clip (ChebyshevStep.py:52)
}

__device__ __forceinline__ void dmma_m8n8k4(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
                 : "+d"(c0), "+d"(c1)
                 : "d"(a), "d"(b));
}

// Internal column layout (feature-major): feature f owns the DC columns f DC .. f DC + DC - 1 holding T_{K0} .. T_{K0 + DC - 1};
// column P = F DC is y.  Two layouts:
//   K0 = 0, DC = D + 1 (D = 0 only): every column of the reference's design matrix, W = P + 1 columns;
//   K0 = 1, DC = D     (D >= 1):      the T_0 columns of the F features are F copies of the same all-ones column
//       (DegreeOptimizer.py:96-119: transforms[0] = ones), so it is kept ONCE, as column P + 1: W = F D + 2 columns instead
//       of F (D + 1) + 1 - at 79 features, degree 3: 239 instead of 317, 10 tiles of the upper triangle instead of 15.  The
//       expand kernel below writes the full (F (D+1) + 1)^2 matrix of the C ABI from it (every entry of a T_0 row / column is
//       the entry of the ones column).
struct GramParams {
    const double* x;       // [n, F]
    const double* y;       // [n]
    double* partial;       // [S][n_tiles][TILE * TILE]
    long long n;
    int F, D, DC, P;       // P = F DC = the y column
    int W;                 // augmented width: P + 1 (+ 1 for the ones column when K0 = 1)
    int T;                 // tiles per edge = ceil(W / TILE)
    int S;                 // sample slices
};

// fill one side's chunk: rows = samples s0 .. s0 + KC, columns c0 .. c0 + TILE of the augmented matrix.
// D1T > 0: the columns per feature DC are the compile-time constant D1T (unrolled recurrence, immediate store offsets).
template <int MAXI, int D1T, int K0>
__device__ __forceinline__ void fill_chunk(const GramParams& p, double* dst, long long s0, int c0, const double (&xv)[MAXI],
                                           int f_lo, int nf) {
    // item = (feature slot j, sample kk) with kk = t % 32 fastest: thread t handles slots t / 32, t / 32 + 4, ...;
    // the values were prefetched into xv[]
    const int D1 = D1T > 0 ? D1T : p.DC;
    const int kk = threadIdx.x & (KC - 1);
    const double lv = s0 + kk < p.n ? 1.0 : 0.0;         // rows beyond the slice are zero rows (last chunk only)
#pragma unroll
    for (int it = 0; it < MAXI; ++it) {
        const int j = (threadIdx.x >> 5) + it * (GRAM_THREADS / KC);
        if (j > nf) break;
        if (j == nf) {                                   // the y column, the ones column (if this tile holds them), the padding
            const int cy = p.P - c0;
            if (cy >= 0 && cy < TILE) dst[cy * LDK + kk] = lv * xv[it];
            if (K0 && cy + 1 >= 0 && cy + 1 < TILE) dst[(cy + 1) * LDK + kk] = lv;
            for (int c = (cy + 1 + K0 >= 0 ? cy + 1 + K0 : 0); c < TILE; ++c)
                if (c0 + c >= p.W) dst[c * LDK + kk] = 0.0;
            continue;
        }
        const int f = f_lo + j;
        const double xc = clip_unit(xv[it]);
        const int cbase = f * D1 - c0;                   // tile column of the feature's first column
        double* q = dst + cbase * LDK + kk;
        double t0 = lv, t1 = lv * xc;                    // T_0, T_1 (times the row's 0 / 1)
        if (K0) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }      // first column = T_1
        if (cbase >= 0 && cbase + D1 <= TILE && f < p.F) {       // the feature's columns lie inside the tile: no checks
            q[0] = t0;
#pragma unroll
            for (int k = 1; k < (D1T > 0 ? D1T : 1); ++k) {
                q[k * LDK] = t1;
                const double t2 = 2.0 * xc * t1 - t0;
                t0 = t1; t1 = t2;
            }
            if (D1T == 0)
                for (int k = 1; k < D1; ++k) {
                    q[k * LDK] = t1;
                    const double t2 = 2.0 * xc * t1 - t0;
                    t0 = t1; t1 = t2;
                }
        } else if (f < p.F) {                            // a feature cut by the tile edge
            for (int k = 0; k < D1; ++k) {
                const int c = cbase + k;
                if (c >= 0 && c < TILE) q[k * LDK] = k == 0 ? t0 : t1;
                if (k >= 1) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }
            }
        }
    }
}

// Straight-line form of fill_chunk: every slot of the thread is evaluated with predicated stores and no branch (the branchy
// form above spent a fifth of the kernel's warp samples on BRA / BSSY / BSYNC, profiles/r02x_ncu_gram.txt).  The columns
// beyond P are zeroed once per launch instead of once per chunk (nothing else ever writes them).
#ifndef GRAM_FLAT_FILL
#define GRAM_FLAT_FILL 1
#endif
template <int MAXI, int D1T, int K0>
__device__ __forceinline__ void fill_chunk_flat(const GramParams& p, double* dst, long long s0, int c0, const double (&xv)[MAXI],
                                                int f_lo, int nf) {
    const int D1 = D1T > 0 ? D1T : p.DC;
    const int kk = threadIdx.x & (KC - 1), jrow = threadIdx.x >> 5;
    const double lv = s0 + kk < p.n ? 1.0 : 0.0;         // rows beyond the slice are zero rows (last chunk only)
#pragma unroll
    for (int it = 0; it < MAXI; ++it) {
        const int j = jrow + it * (GRAM_THREADS / KC);
        const bool isy = j == nf, feat = j < nf;         // slot nf is the y column; slots beyond it do nothing
        const double v = xv[it];
        const int cbase = isy ? p.P - c0 : (f_lo + j) * D1 - c0;
        double* q = dst + cbase * LDK + kk;
        const double xc = clip_unit(v);
        double t0 = lv, t1 = lv * xc;
        if (K0) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }      // first column = T_1
        if ((feat || isy) && (unsigned)cbase < (unsigned)TILE) q[0] = isy ? lv * v : t0;
        if (K0 && isy && (unsigned)(cbase + 1) < (unsigned)TILE) q[LDK] = lv;    // the ones column follows y
        if constexpr (D1T > 0) {
#pragma unroll
            for (int k = 1; k < D1T; ++k) {
                if (feat && (unsigned)(cbase + k) < (unsigned)TILE) q[k * LDK] = t1;
                const double t2 = 2.0 * xc * t1 - t0;
                t0 = t1; t1 = t2;
            }
        } else {
            for (int k = 1; k < D1; ++k) {
                if (feat && (unsigned)(cbase + k) < (unsigned)TILE) q[k * LDK] = t1;
                const double t2 = 2.0 * xc * t1 - t0;
                t0 = t1; t1 = t2;
            }
        }
    }
}

// Rejected variants (correct, A/B on one box, profiles/README.md): double-buffered chunks with one barrier (round 1),
// a producer warp + mbarrier ring, cp.async input staging (round 2, first session), and chunks double buffered with the
// production of chunk i + 1 sliced between the eight DMMA groups of chunk i (3 CTAs per SM: 3.55 ms against 3.49 ms);
// 5 CTAs per SM at 96 registers (164 bytes spilled): 3.93 ms.
// MAXI = most feature slots a thread fills per chunk and side: (64 / DC + 2 features + the y slot) / 4
template <int MAXI, int D1T, int K0, int MINB = (MAXI > 9 ? 2 : (MAXI > 6 ? 3 : 4))>
__global__ void __launch_bounds__(GRAM_THREADS, MINB) qkan_cheb_gram_kernel(const GramParams p) {
    __shared__ __align__(16) double As[TILE * LDK];
    __shared__ __align__(16) double Bs[TILE * LDK];
    // upper-triangle tile (ti <= tj) from the linear tile index
    int ti = 0, rem = blockIdx.x;
    while (rem >= p.T - ti) { rem -= p.T - ti; ++ti; }
    const int tj = ti + rem;
    const bool diag = ti == tj;
    const int ca = ti * TILE, cb = tj * TILE;
    const int D1 = p.DC;
    // features touched by each side's 64 columns (+ one pseudo feature slot per sample: y / ones / padding)
    const int fa_lo = ca / D1, fb_lo = cb / D1;
    auto nfeat = [&](int c0, int f_lo) {
        int c_hi = c0 + TILE - 1;
        if (c_hi > p.P - 1) c_hi = p.P - 1;
        const int n = c_hi >= c0 ? c_hi / D1 - f_lo + 1 : 0;
        return n;
    };
    const int nfa = nfeat(ca, fa_lo), nfb = nfeat(cb, fb_lo);
    double xa[MAXI], xb[MAXI];

    const long long per = (p.n + p.S - 1) / p.S;
    const long long s_begin = (long long)blockIdx.y * per;
    long long s_end = s_begin + per;
    if (s_end > p.n) s_end = p.n;

    auto prefetch = [&](long long s0, double (&xv)[MAXI], int f_lo, int nf) {
        const int kk = threadIdx.x & (KC - 1);
#pragma unroll
        for (int it = 0; it < MAXI; ++it) {
            const int j = (threadIdx.x >> 5) + it * (GRAM_THREADS / KC);
            if (j > nf) break;
            double v = 0.0;
            if (s0 + kk < s_end) {
                if (j == nf) v = p.y[s0 + kk];
                else if (f_lo + j < p.F) v = p.x[(s0 + kk) * p.F + f_lo + j];
            }
            xv[it] = v;
        }
    };

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // the warp's 32 x 32 quadrant, rotated from CTA to CTA: warp w of every resident CTA sits on SM sub-partition w, and the
    // strictly lower quadrant of a diagonal tile (the transpose of the upper one) is not computed - without the rotation
    // one sub-partition's tensor pipe would get all the idle warps
    const int quad = (warp + blockIdx.x + blockIdx.y) & 3;
    const int wm = (quad & 1) * 32, wn = (quad >> 1) * 32;
    const bool idle = diag && wm > wn;
    const int lr = lane >> 2, lk = lane & 3;
    double acc[4][4][2];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    // the chunk loop treats samples beyond s_end as zero rows (live test uses p.n; slices end on s_end)
    GramParams q = p;
    q.n = s_end;
    if (s_begin < s_end) {
        prefetch(s_begin, xa, fa_lo, nfa);
        if (!diag) prefetch(s_begin, xb, fb_lo, nfb);
    }
    // compile-time degree (D <= 4): the straight-line producer (774 456 x 79: D = 3 3.50 -> 3.39 ms, D = 1 1.59 -> 1.53 ms); run-time
    // degree: the branchy one (a run-time loop of predicated stores is slower: D = 8, 200 k rows: 4.86 against 5.70 ms)
    constexpr bool flat = GRAM_FLAT_FILL && D1T > 0;
    if constexpr (flat)
        for (int i = threadIdx.x; i < TILE * LDK; i += GRAM_THREADS) As[i] = Bs[i] = 0.0;    // the columns beyond W stay zero
    for (long long s0 = s_begin; s0 < s_end; s0 += KC) {
        __syncthreads();                                 // the previous chunk's fragments are consumed
        if constexpr (flat) {
            fill_chunk_flat<MAXI, D1T, K0>(q, As, s0, ca, xa, fa_lo, nfa);
            if (!diag) fill_chunk_flat<MAXI, D1T, K0>(q, Bs, s0, cb, xb, fb_lo, nfb);
        } else {
            fill_chunk<MAXI, D1T, K0>(q, As, s0, ca, xa, fa_lo, nfa);
            if (!diag) fill_chunk<MAXI, D1T, K0>(q, Bs, s0, cb, xb, fb_lo, nfb);
        }
        __syncthreads();
        if (s0 + KC < s_end) {                           // next chunk's x / y: in flight during the MMAs
            prefetch(s0 + KC, xa, fa_lo, nfa);
            if (!diag) prefetch(s0 + KC, xb, fb_lo, nfb);
        }
        const double* Bsrc = diag ? As : Bs;
        if (idle) continue;
#pragma unroll
        for (int k4 = 0; k4 < KC; k4 += 4) {
            double a[4], b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) a[i] = As[(wm + i * 8 + lr) * LDK + k4 + lk];      // A[m = lane/4][k = lane%4]
#pragma unroll
            for (int j = 0; j < 4; ++j) b[j] = Bsrc[(wn + j * 8 + lr) * LDK + k4 + lk];    // B[k = lane%4][n = lane/4]
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) dmma_m8n8k4(acc[i][j][0], acc[i][j][1], a[i], b[j]);
        }
    }
    // C fragment: (row = lane/4, cols 2 (lane%4) + {0, 1}) of each 8 x 8 block
    double* out = p.partial + ((size_t)blockIdx.y * gridDim.x + blockIdx.x) * (TILE * TILE);
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const int r = wm + i * 8 + lr, c = wn + j * 8 + 2 * lk;
            out[r * TILE + c] = acc[i][j][0];
            out[r * TILE + c + 1] = acc[i][j][1];
        }
}

// sum the slices in a fixed order into the W x W Gram matrix of the internal column layout (both triangles).
// grid (tiles, TILE * TILE / blockDim): one element per thread, the slices read as coalesced 2 KB rows (a 15-CTA version that
// walked a whole tile per CTA took 0.4 ms of the 4.07 ms Gram step).  Diagonal tiles contribute their upper triangle only
// (the Gram kernel does not compute their strictly lower quadrant).
__global__ void __launch_bounds__(256) qkan_cheb_gram_reduce_kernel(const double* __restrict__ partial, int n_tiles, int S, int T, int W,
                                                                     double* __restrict__ Gw) {
    const int tile = blockIdx.x;
    int ti = 0, rem = tile;
    while (rem >= T - ti) { rem -= T - ti; ++ti; }
    const int tj = ti + rem;
    const int e = blockIdx.y * blockDim.x + threadIdx.x;
    const int r = ti * TILE + e / TILE, c = tj * TILE + e % TILE;
    if (r >= W || c >= W || (ti == tj && r > c)) return;
    const double* src = partial + (size_t)tile * (TILE * TILE) + e;
    const size_t stride = (size_t)n_tiles * (TILE * TILE);
    double s = 0.0;
    int q = 0;
    for (; q + 4 <= S; q += 4) {                             // four loads in flight; the order of the additions is fixed
        const double v0 = src[(size_t)q * stride], v1 = src[(size_t)(q + 1) * stride];
        const double v2 = src[(size_t)(q + 2) * stride], v3 = src[(size_t)(q + 3) * stride];
        s += v0; s += v1; s += v2; s += v3;
    }
    for (; q < S; ++q) s += src[(size_t)q * stride];
    Gw[(size_t)r * W + c] = s;
    Gw[(size_t)c * W + r] = s;
}

// the (F (D+1) + 1)^2 result of the C ABI, columns in the reference's np.hstack order (k F + f, then y), from the internal
// matrix: a T_k column with k < K0 is the ones column, the others sit at f DC + k - K0
__global__ void __launch_bounds__(256) qkan_cheb_gram_expand_kernel(const double* __restrict__ Gw, int W, int F, int D, int DC, int K0,
                                                                     double* __restrict__ G) {
    const int PF = F * (D + 1);                              // the y row / column of the result
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (long long)(PF + 1) * (PF + 1)) return;
    const int rr = (int)(i / (PF + 1)), cc = (int)(i % (PF + 1));
    auto internal = [&](int c) {
        if (c == PF) return F * DC;
        const int k = c / F, f = c - k * F;
        return k < K0 ? F * DC + 1 : f * DC + (k - K0);
    };
    G[i] = Gw[(size_t)internal(rr) * W + internal(cc)];
}

// ---- residual pass: one warp per sample, lanes over features
// coef [D+1][P] degree-major (zero where the column's degree exceeds d).  Per CTA partial sums:
//   sums[cta][d][0] = sum r_d^2, [1] = sum w r_d^2;   tail[cta] = {sum (y - ybar)^2, sum w y^2, sum w, sum y}
//   xtr[cta][d][P]  = X_D^T r_d   (optional)
constexpr int RES_THREADS = 256;

// JF > 0: F <= 32 JF and the X^T r accumulators of a lane (its JF features x D1 degrees x D1 fits) live in registers for
// the whole launch, flushed once per warp at the end; JF = 0: any F, accumulation by shared-memory atomics per sample.
template <int D1, int JF>
__global__ void __launch_bounds__(RES_THREADS) qkan_cheb_residual_kernel(const double* x, const double* y, const double* w, long long n, int F,
                                                                        const double* coef, double ybar, double* sums,
                                                                        double* tail, double* xtr) {
    extern __shared__ double sm[];
    const int P = F * D1;
    double* s_xtr = sm;                                  // [D1][P] CTA accumulators of X^T r_d (if requested)
    double* s_red = sm + (xtr ? (size_t)D1 * P : 0);     // [warps][2 D1 + 4]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = RES_THREADS / 32;
    if (xtr)
        for (int i = threadIdx.x; i < D1 * P; i += RES_THREADS) s_xtr[i] = 0.0;
    __syncthreads();
    double a_sse[D1], a_wsse[D1];
#pragma unroll
    for (int d = 0; d < D1; ++d) a_sse[d] = a_wsse[d] = 0.0;
    double a_tot = 0.0, a_wyy = 0.0, a_w = 0.0, a_y = 0.0;
    constexpr int JR = JF > 0 ? JF : 1;
    double a_x[JR][D1][D1];                              // [feature slot][degree k][fit d]
#pragma unroll
    for (int jj = 0; jj < JR; ++jj)
#pragma unroll
        for (int k = 0; k < D1; ++k)
#pragma unroll
            for (int d = 0; d < D1; ++d) a_x[jj][k][d] = 0.0;
    const long long gw = (long long)blockIdx.x * nwarp + warp, nw = (long long)gridDim.x * nwarp;
    // register path: the next sample's inputs are loaded while the current one is evaluated (the pass is latency bound)
    double xn[JR], yn = 0.0, wn = 1.0;
    auto load_sample = [&](long long s) {
        if constexpr (JF > 0) {
#pragma unroll
            for (int jj = 0; jj < JF; ++jj) {
                const int f = lane + 32 * jj;
                xn[jj] = (s < n && f < F) ? x[s * F + f] : 0.0;
            }
            yn = s < n ? y[s] : 0.0;
            wn = (s < n && w) ? w[s] : 1.0;
        }
    };
    load_sample(gw);
    for (long long s = gw; s < n; s += nw) {
        double pred[D1];
#pragma unroll
        for (int d = 0; d < D1; ++d) pred[d] = 0.0;
        double tv[JR][D1];                               // T_k of the lane's features (register path)
        double xcur[JR];
#pragma unroll
        for (int jj = 0; jj < JR; ++jj) xcur[jj] = xn[jj];
        const double ycur = yn, wcur = wn;
        load_sample(s + nw);
        if constexpr (JF > 0) {
#pragma unroll
            for (int jj = 0; jj < JF; ++jj) {
                const int f = lane + 32 * jj;
                const bool in = f < F;
                const double xc = in ? clip_unit(xcur[jj]) : 0.0;
                double t0 = in ? 1.0 : 0.0, t1 = in ? xc : 0.0;
#pragma unroll
                for (int k = 0; k < D1; ++k) {
                    const double tk = k == 0 ? t0 : t1;
                    tv[jj][k] = tk;
#pragma unroll
                    for (int d = k; d < D1; ++d) pred[d] = fma(tk, in ? coef[(size_t)d * P + k * F + f] : 0.0, pred[d]);
                    if (k >= 1) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }
                }
            }
        } else {
            for (int f = lane; f < F; f += 32) {
                const double xc = clip_unit(x[s * F + f]);
                double t0 = 1.0, t1 = xc;
#pragma unroll
                for (int k = 0; k < D1; ++k) {
                    const double tk = k == 0 ? 1.0 : t1;
#pragma unroll
                    for (int d = k; d < D1; ++d) pred[d] = fma(tk, coef[(size_t)d * P + k * F + f], pred[d]);
                    if (k >= 1) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }
                }
            }
        }
#pragma unroll
        for (int d = 0; d < D1; ++d)
            for (int m = 16; m >= 1; m >>= 1) pred[d] += __shfl_xor_sync(0xffffffffu, pred[d], m);
        const double yv = JF > 0 ? ycur : y[s], wv = JF > 0 ? wcur : (w ? w[s] : 1.0);
        double r[D1];
#pragma unroll
        for (int d = 0; d < D1; ++d) r[d] = yv - pred[d];
        if (lane == 0) {
#pragma unroll
            for (int d = 0; d < D1; ++d) { a_sse[d] += r[d] * r[d]; a_wsse[d] += wv * r[d] * r[d]; }
            a_tot += (yv - ybar) * (yv - ybar);
            a_wyy += wv * yv * yv;
            a_w += wv;
            a_y += yv;
        }
        if (xtr) {
            if constexpr (JF > 0) {
#pragma unroll
                for (int jj = 0; jj < JF; ++jj)
#pragma unroll
                    for (int k = 0; k < D1; ++k)
#pragma unroll
                        for (int d = 0; d < D1; ++d) a_x[jj][k][d] = fma(tv[jj][k], r[d], a_x[jj][k][d]);
            } else {
                for (int f = lane; f < F; f += 32) {
                    const double xc = clip_unit(x[s * F + f]);
                    double t0 = 1.0, t1 = xc;
#pragma unroll
                    for (int k = 0; k < D1; ++k) {
                        const double tk = k == 0 ? 1.0 : t1;
#pragma unroll
                        for (int d = 0; d < D1; ++d) atomicAdd(&s_xtr[(size_t)d * P + k * F + f], tk * r[d]);
                        if (k >= 1) { const double t2 = 2.0 * xc * t1 - t0; t0 = t1; t1 = t2; }
                    }
                }
            }
        }
    }
    if constexpr (JF > 0) {
        if (xtr) {                                       // one flush per warp
#pragma unroll
            for (int jj = 0; jj < JF; ++jj) {
                const int f = lane + 32 * jj;
                if (f < F) {
#pragma unroll
                    for (int k = 0; k < D1; ++k)
#pragma unroll
                        for (int d = 0; d < D1; ++d) atomicAdd(&s_xtr[(size_t)d * P + k * F + f], a_x[jj][k][d]);
                }
            }
        }
    }
    constexpr int Q = 2 * D1 + 4;
    if (lane == 0) {
        double* o = s_red + warp * Q;
#pragma unroll
        for (int d = 0; d < D1; ++d) { o[2 * d] = a_sse[d]; o[2 * d + 1] = a_wsse[d]; }
        o[2 * D1] = a_tot; o[2 * D1 + 1] = a_wyy; o[2 * D1 + 2] = a_w; o[2 * D1 + 3] = a_y;
    }
    __syncthreads();
    if (threadIdx.x < Q) {
        double sacc = 0.0;
        for (int wq = 0; wq < nwarp; ++wq) sacc += s_red[wq * Q + threadIdx.x];
        if (threadIdx.x < 2 * D1) sums[(size_t)blockIdx.x * 2 * D1 + threadIdx.x] = sacc;
        else tail[(size_t)blockIdx.x * 4 + threadIdx.x - 2 * D1] = sacc;
    }
    if (xtr)
        for (int i = threadIdx.x; i < D1 * P; i += RES_THREADS) xtr[(size_t)blockIdx.x * D1 * P + i] = s_xtr[i];
}

// ---- residual pass, tile form (D <= 4, F <= 128: the reference's workload, 79 features at max_degree 3)
// The warp-per-sample kernel above spends 390 of its 540 instructions per sample on coefficient addresses / loads and on
// the 5-step butterflies of the D + 1 predictions (profiles/r02Z_ncu_degree_kernels.txt: 228 registers, one CTA per SM,
// FP64 pipe 22 % busy, 0.66 ms at 774 456 x 79 where the x rows alone stream from HBM in 0.07 ms).  Here a CTA walks
// tiles of 64 samples, staged by 1-D TMA bulk copies (two tiles in flight), in two phases with different lane mappings:
//   phase 1 (predictions): four lanes per sample, each over a contiguous quarter of the features; the coefficients of a
//       feature are one 16-byte-aligned run in shared memory (immediate offsets, no index arithmetic), the T_0 columns
//       are summed once per launch (they do not depend on x), and the four partial predictions meet in two xor steps;
//   phase 2 (X^T r, only when requested): a warp per sample, a lane per feature as above, with the residuals of the
//       sample broadcast from shared memory; the accumulators stay in registers for the whole launch.
// Same outputs (per-CTA partial sums in the same layout) as the warp-per-sample kernel.
// Measured and not kept: 64-sample tiles (SPL = 1) at 128 registers so that two CTAs share an SM and their passes overlap -
// 0.286 / 0.204 ms against 0.294 / 0.194 ms (profiles/r03g_bench_residuals_two_ctas_rejected.jsonl): what the overlap gains, the
// coefficient loads serving one sample instead of two lose.
constexpr int RT_THREADS = 256;
constexpr int RT_TS1 = 64;                                  // samples per tile and per phase-1 round (8 warps x 8 samples); tiles hold
                                                            // SPL rounds, SPL = 2 when two tiles of 128 samples fit shared memory

__host__ __device__ constexpr int rt_tri(int D1) { return D1 * (D1 - 1) / 2; }           // (k, d) pairs with 1 <= k <= d < D1
__host__ __device__ constexpr int rt_cw(int D1) { return (rt_tri(D1) + 1) & ~1; }        // doubles per feature (even: LDS.128)
__host__ __device__ constexpr int rt_idx(int D1, int k, int d) {                         // position of (k, d) in a feature's run
    int i = 0;
    for (int kk = 1; kk < k; ++kk) i += D1 - kk;
    return i + (d - k);
}
// The coefficient runs of the four feature quarters start 8 banks apart (quarter stride = 4 mod 16 doubles): the 16-byte loads of
// a quarter warp - two samples x four quarters - then touch four different bank groups (at F = 79, D = 3 the unpadded stride put
// quarters 0 / 2 and 1 / 3 on the same banks: every coefficient load took two passes).
__host__ __device__ inline int rt_qstride(int F, int D1) {
    const int raw = ((F + 3) >> 2) * rt_cw(D1);
    return raw == 0 ? 0 : raw + ((4 - raw % 16) + 16) % 16;
}
// shared memory in doubles: xs[2][TS F] | cf[4 quarter strides] | c0[D1P] | rs[TS D1P] | s_xtr[D1 P] | s_red[warps][3 D1 + 4] | mbar[2]
struct ResTileSmem {
    int tile, cf, c0, rs, xtr, red, mbar, total;
};
__host__ __device__ inline ResTileSmem rt_smem(int F, int D1, bool want_xtr, int TS) {
    ResTileSmem s;
    const int D1P = (D1 + 1) & ~1;
    s.tile = TS * F;
    s.cf = 2 * s.tile;
    s.c0 = s.cf + 4 * rt_qstride(F, D1);
    s.rs = s.c0 + D1P;
    s.xtr = s.rs + TS * D1P;
    s.red = s.xtr + (want_xtr ? ((D1 * D1 * F + 1) & ~1) : 0);
    s.mbar = s.red + (((RT_THREADS / 32) * (3 * D1 + 4) + 1) & ~1);
    s.total = s.mbar + 2;
    return s;
}

__device__ __forceinline__ unsigned rt_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void rt_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void rt_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(rt_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void rt_mbar_wait(unsigned long long* bar, unsigned parity) {
    const unsigned addr = rt_smem_u32(bar);
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    }
}
// 1-D TMA bulk copy global -> shared, completion on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void rt_tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(rt_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(rt_smem_u32(bar))
                 : "memory");
}

template <int D1, int JF, int SPL>
__global__ void __launch_bounds__(RT_THREADS, 1) qkan_cheb_residual_tile_kernel(const double* __restrict__ x, const double* __restrict__ y,
                                                                               const double* __restrict__ w, long long n, int F,
                                                                               const double* __restrict__ coef, double ybar, double* sums,
                                                                               double* tail, double* xtr, int tma_ok) {
    extern __shared__ __align__(128) double rt_sm[];
    constexpr int TS = RT_TS1 * SPL, NT = RT_THREADS, NW = NT / 32;
    constexpr int CW = rt_cw(D1), CWA = CW > 0 ? CW : 2, D1P = (D1 + 1) & ~1;
    const int P = F * D1;
    const ResTileSmem L = rt_smem(F, D1, xtr != nullptr, TS);
    const int FQ = (F + 3) >> 2, QS = rt_qstride(F, D1);     // features per quarter, its stride in cf
    double* const xs = rt_sm;
    double* const cf = rt_sm + L.cf;
    double* const c0 = rt_sm + L.c0;
    double* const rs = rt_sm + L.rs;
    double* const s_xtr = rt_sm + L.xtr;
    double* const s_red = rt_sm + L.red;
    unsigned long long* const mbar = reinterpret_cast<unsigned long long*>(rt_sm + L.mbar);
    const int tid = threadIdx.x, lane = tid & 31, wq = tid >> 5;

    if (tid == 0) {
        rt_mbar_init(&mbar[0], 1);
        rt_mbar_init(&mbar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // a feature's coefficients of degree k >= 1 as one run: cf[(f / FQ) QS + (f % FQ) CW + rt_idx(k, d)] = coef[d][k F + f], d >= k
    if constexpr (CW > 0) {
        for (int i = tid; i < F * CW; i += NT) {
            const int f = i / CW;
            int rem = i - f * CW, k = 1;
            while (k < D1 && rem >= D1 - k) { rem -= D1 - k; ++k; }
            const int fq = f / FQ;
            cf[fq * QS + (f - fq * FQ) * CW + (i - f * CW)] = k < D1 ? coef[(size_t)(k + rem) * P + (size_t)k * F + f] : 0.0;
        }
    }
    // the T_0 columns do not depend on x: their coefficients are summed once (warp d sums fit d)
    for (int d = wq; d < D1P; d += NW) {
        double s = 0.0;
        if (d < D1)
            for (int f = lane; f < F; f += 32) s += coef[(size_t)d * P + f];
        for (int m = 16; m >= 1; m >>= 1) s += __shfl_xor_sync(0xffffffffu, s, m);
        if (lane == 0) c0[d] = s;
    }
    if (xtr)
        for (int i = tid; i < D1 * P; i += NT) s_xtr[i] = 0.0;

    const long long n_tiles = (n + TS - 1) / TS;
    auto tile_n = [&](long long t) -> int {
        const long long left = n - t * TS;
        return left < TS ? (int)left : TS;
    };
    auto by_tma = [&](long long t) -> bool { return tma_ok && ((tile_n(t) * F) & 1) == 0; };
    auto issue = [&](long long t, int b) {
        const int cnt = tile_n(t) * F;
        const double* src = x + (size_t)t * L.tile;
        double* dst = xs + (size_t)b * L.tile;
        if (by_tma(t)) {
            if (tid == 0) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                rt_mbar_expect_tx(&mbar[b], (unsigned)cnt * 8u);
                rt_tma_load_1d(dst, src, (unsigned)cnt * 8u, &mbar[b]);
            }
        } else {
            for (int q = tid; q < cnt; q += NT) dst[q] = src[q];
        }
    };
    __syncthreads();                                         // barriers initialised, tables complete

    double a_sse[D1], a_wsse[D1], a_r[D1];
#pragma unroll
    for (int d = 0; d < D1; ++d) a_sse[d] = a_wsse[d] = a_r[d] = 0.0;
    double a_tot = 0.0, a_wyy = 0.0, a_w = 0.0, a_y = 0.0;
    // X^T r accumulators of the lane's features for the degrees k >= 1 (the T_0 rows are sum r_d for every feature: a_r)
    constexpr int KX = D1 > 1 ? D1 - 1 : 1;
    double a_x[JF][KX][D1];                                  // [feature slot][degree k - 1][fit d]
#pragma unroll
    for (int jj = 0; jj < JF; ++jj)
#pragma unroll
        for (int k = 0; k < KX; ++k)
#pragma unroll
            for (int d = 0; d < D1; ++d) a_x[jj][k][d] = 0.0;

    const long long step = gridDim.x;
    long long t = blockIdx.x;
    if (t < n_tiles) issue(t, 0);
    if (t + step < n_tiles) issue(t + step, 1);
    __syncthreads();                                         // (tiles staged by plain loads)
    unsigned ph0 = 0, ph1 = 0;
    int b = 0;
    const int sl = lane >> 2, q4 = lane & 3;
    const int st = wq * 8 + sl;                              // phase 1: the lane's sample of each round of 64
    const int f_lo = q4 * FQ < F ? q4 * FQ : F, f_hi = f_lo + FQ < F ? f_lo + FQ : F;
    const double* const cfq = cf + q4 * QS - f_lo * CW;      // + f CW = the run of feature f of this lane's quarter
    const bool want_x = xtr != nullptr && D1 > 1;
    for (; t < n_tiles; t += step, b ^= 1) {
        const int nsamp = tile_n(t);
        if (by_tma(t)) {
            if (b == 0) { rt_mbar_wait(&mbar[0], ph0); ph0 ^= 1; }
            else        { rt_mbar_wait(&mbar[1], ph1); ph1 ^= 1; }
        }
        double* const xt = xs + (size_t)b * L.tile;
        if constexpr (D1 > 1) {
            // ---- clip pass (np.clip, ChebyshevStep.py:52): only inputs with |v| >= 1 (or NaN) can change, decided on the high
            // words with integer instructions, two inputs per step; the phases below then read clipped inputs
            const int cnt = nsamp * F;
            double2* const x2p = reinterpret_cast<double2*>(xt);
            for (int i = tid; i < (cnt >> 1); i += NT) {
                const double2 v = x2p[i];
                const unsigned h0 = (unsigned)__double2hiint(v.x) & 0x7fffffffu, h1 = (unsigned)__double2hiint(v.y) & 0x7fffffffu;
                if ((h0 > h1 ? h0 : h1) >= 0x3ff00000u) x2p[i] = make_double2(clip_unit(v.x), clip_unit(v.y));
            }
            if ((cnt & 1) && tid == 0) xt[cnt - 1] = clip_unit(xt[cnt - 1]);
            __syncthreads();
        }
        {   // ---- phase 1: predictions of the D + 1 fits for the lane's SPL samples, residuals, the sums of the scores
            bool valid[SPL];
            double yv[SPL], wv[SPL], pred[SPL][D1];
#pragma unroll
            for (int sp = 0; sp < SPL; ++sp) {
                valid[sp] = st + sp * RT_TS1 < nsamp;
                const long long s = t * TS + st + sp * RT_TS1;
                yv[sp] = 0.0; wv[sp] = 1.0;
                if (valid[sp]) {
                    yv[sp] = y[s];
                    if (w) wv[sp] = w[s];
                }
#pragma unroll
                for (int d = 0; d < D1; ++d) pred[sp][d] = q4 == 0 ? c0[d] : 0.0;
            }
            if constexpr (D1 > 1) {
                const double* xr = xt + st * F;
                for (int f = f_lo; f < f_hi; ++f) {
                    double cw[CWA];
                    const double2* cp = reinterpret_cast<const double2*>(cfq + f * CW);
#pragma unroll
                    for (int i = 0; i < CW / 2; ++i) {
                        const double2 v = cp[i];
                        cw[2 * i] = v.x;
                        cw[2 * i + 1] = v.y;
                    }
#pragma unroll
                    for (int sp = 0; sp < SPL; ++sp) {
                        const double xc = xr[sp * RT_TS1 * F + f];
                        const double x2 = xc + xc;
                        double t0 = 1.0, t1 = xc;
#pragma unroll
                        for (int k = 1; k < D1; ++k) {
#pragma unroll
                            for (int d = k; d < D1; ++d) pred[sp][d] = fma(t1, cw[rt_idx(D1, k, d)], pred[sp][d]);
                            const double t2 = fma(x2, t1, -t0);
                            t0 = t1; t1 = t2;
                        }
                    }
                }
            }
#pragma unroll
            for (int sp = 0; sp < SPL; ++sp) {
#pragma unroll
                for (int d = 0; d < D1; ++d) {
                    pred[sp][d] += __shfl_xor_sync(0xffffffffu, pred[sp][d], 1);
                    pred[sp][d] += __shfl_xor_sync(0xffffffffu, pred[sp][d], 2);
                }
                if (valid[sp] && q4 == 0) {
#pragma unroll
                    for (int d = 0; d < D1; ++d) {
                        const double r = yv[sp] - pred[sp][d];
                        a_sse[d] = fma(r, r, a_sse[d]);
                        a_wsse[d] = fma(wv[sp] * r, r, a_wsse[d]);
                        a_r[d] += r;
                        if (want_x) rs[(st + sp * RT_TS1) * D1P + d] = r;
                    }
                    a_tot = fma(yv[sp] - ybar, yv[sp] - ybar, a_tot);
                    a_wyy = fma(wv[sp] * yv[sp], yv[sp], a_wyy);
                    a_w += wv[sp];
                    a_y += yv[sp];
                }
            }
        }
        __syncthreads();                                     // residuals of the tile in rs; phase 1 done with xt
        if constexpr (D1 > 1) {
            if (want_x) {   // ---- phase 2: X^T r of the tile for the degrees k >= 1 (warp per sample, lane per feature)
                for (int s2 = wq; s2 < nsamp; s2 += NW) {
                    const double* xr = xt + s2 * F;
                    double r[D1P];
                    const double2* rp = reinterpret_cast<const double2*>(rs + s2 * D1P);
#pragma unroll
                    for (int i = 0; i < D1P / 2; ++i) {
                        const double2 v = rp[i];
                        r[2 * i] = v.x;
                        r[2 * i + 1] = v.y;
                    }
#pragma unroll
                    for (int jj = 0; jj < JF; ++jj) {
                        // a lane past the row (f >= F) reads the next row / the tables: its sums are never flushed
                        const double xc = xr[lane + 32 * jj];
                        const double x2 = xc + xc;
                        double t0 = 1.0, t1 = xc;
#pragma unroll
                        for (int k = 1; k < D1; ++k) {
#pragma unroll
                            for (int d = 0; d < D1; ++d) a_x[jj][k - 1][d] = fma(t1, r[d], a_x[jj][k - 1][d]);
                            const double t2 = fma(x2, t1, -t0);
                            t0 = t1; t1 = t2;
                        }
                    }
                }
                __syncthreads();                             // rs and xt free again
            }
        }
        if (t + 2 * step < n_tiles) issue(t + 2 * step, b);  // lands while the next tile is evaluated
    }

    if constexpr (D1 > 1) {
        if (want_x) {                                        // one flush per warp
#pragma unroll
            for (int jj = 0; jj < JF; ++jj) {
                const int f = lane + 32 * jj;
                if (f < F) {
#pragma unroll
                    for (int k = 1; k < D1; ++k)
#pragma unroll
                        for (int d = 0; d < D1; ++d) atomicAdd(&s_xtr[(size_t)d * P + k * F + f], a_x[jj][k - 1][d]);
                }
            }
        }
    }
    constexpr int Q = 2 * D1 + 4, QR = Q + D1;               // the outputs' sums, then sum r_d
    {
        double v[QR];
#pragma unroll
        for (int d = 0; d < D1; ++d) { v[2 * d] = a_sse[d]; v[2 * d + 1] = a_wsse[d]; v[Q + d] = a_r[d]; }
        v[2 * D1] = a_tot; v[2 * D1 + 1] = a_wyy; v[2 * D1 + 2] = a_w; v[2 * D1 + 3] = a_y;
#pragma unroll
        for (int i = 0; i < QR; ++i) {
            for (int m = 16; m >= 1; m >>= 1) v[i] += __shfl_xor_sync(0xffffffffu, v[i], m);
            if (lane == 0) s_red[wq * QR + i] = v[i];
        }
    }
    __syncthreads();
    if (tid < QR) {
        double sacc = 0.0;
        for (int wi = 0; wi < NW; ++wi) sacc += s_red[wi * QR + tid];
        if (tid < 2 * D1) sums[(size_t)blockIdx.x * 2 * D1 + tid] = sacc;
        else if (tid < Q) tail[(size_t)blockIdx.x * 4 + tid - 2 * D1] = sacc;
        else if (xtr) {                                      // the T_0 rows of X^T r_d: sum r_d for every feature
            const int d = tid - Q;
            for (int f = 0; f < F; ++f) s_xtr[(size_t)d * P + f] = sacc;
        }
    }
    __syncthreads();
    if (xtr)
        for (int i = tid; i < D1 * P; i += NT) xtr[(size_t)blockIdx.x * D1 * P + i] = s_xtr[i];
}

__global__ void qkan_cheb_features_kernel(const double* x, long long n, int F, int D, double* out) {
    // out [D+1][n][F]: the reference's transforms[d] (DegreeOptimizer.py:96-119)
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * F) return;
    const double xc = clip_unit(x[i]);
    double t0 = 1.0, t1 = xc;
    out[i] = 1.0;
    for (int k = 1; k <= D; ++k) {
        out[(size_t)k * n * F + i] = t1;
        const double t2 = 2.0 * xc * t1 - t0;
        t0 = t1; t1 = t2;
    }
}

int fail(int code, const char* msg) {
    qkan_set_last_error(msg);
    return code;
}
int cuda_fail(cudaError_t e, const char* what) {
    char buf[256];
    snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    qkan_set_last_error(buf);
    return QKAN_ERR_CUDA;
}

}  // namespace

namespace {
// internal layout of the Gram kernels for degree D (see GramParams)
struct GramLayout {
    int K0, DC, P, W, T, n_tiles;
};
// resident CTAs per SM the Gram kernel is compiled for.  D = 3: three CTAs at 168 registers instead of four at 128 - under the
// 128-register cap the straight-line producer recomputed the clip and the row's 0 / 1 several times per slot (583 instructions
// per chunk between the two barriers; 330 at 168 registers): 2.71 -> 2.42 ms at 774 456 x 79 (profiles/r03h_gram_minb.jsonl).
// QKAN_GRAM_MINB = 3 | 4 overrides it for D >= 3 (A/B aid).
int gram_minb(int D) {
    if (D < 3) return D == 2 ? 3 : 2;
    int m = D == 3 ? 3 : 4;
    if (const char* e = getenv("QKAN_GRAM_MINB")) {
        const int v = atoi(e);
        if (v == 3 || v == 4) m = v;
    }
    return m;
}
GramLayout gram_layout(int F, int D) {
    GramLayout g;
    g.K0 = D >= 1 ? 1 : 0;
    g.DC = D >= 1 ? D : 1;
    g.P = F * g.DC;
    g.W = g.P + 1 + g.K0;
    g.T = (g.W + TILE - 1) / TILE;
    g.n_tiles = g.T * (g.T + 1) / 2;
    return g;
}
int gram_slices(int64_t n, int D, const GramLayout& g) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // eight whole waves of CTAs and not one CTA more (rounding up - 2 370 CTAs on 2 368 resident slots at 774 456 x 79,
    // D = 3 - left the SMs idle 15 % of the kernel while two stragglers ran a fifth wave; 4 waves 3.68 ms, 6: 3.55, 8: 3.49,
    // 12: 3.49), at least 8 chunks of samples per slice
    const int per_sm = gram_minb(D);                     // the launch bounds of qkan_cheb_gram_kernel<MAXI, ., ., MINB>
    int waves = 8;
    if (const char* e = getenv("QKAN_GRAM_WAVES")) waves = atoi(e) > 0 ? atoi(e) : waves;   // tuning aid
    int S = waves * per_sm * sms / g.n_tiles;
    const long long max_s = n / (8 * KC) > 0 ? n / (8 * KC) : 1;
    if (S > max_s) S = (int)max_s;
    if (S < 1) S = 1;
    return S;
}
}  // namespace

extern "C" int qkan_cheb_gram_workspace(int64_t n, int F, int D, int64_t* bytes, int* slices) {
    if (!bytes || n < 0 || F < 1 || D < 0 || D > MAX_D) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_gram_workspace: bad arguments");
    const GramLayout g = gram_layout(F, D);
    const int S = gram_slices(n, D, g);
    // the slices' partial tiles, then the W x W matrix of the internal layout
    *bytes = ((int64_t)S * g.n_tiles * TILE * TILE + (int64_t)g.W * g.W) * (int64_t)sizeof(double);
    if (slices) *slices = S;
    return QKAN_OK;
}

extern "C" int qkan_cheb_gram(const double* x, const double* y, int64_t n, int F, int D, double* G, void* workspace,
                              int64_t workspace_bytes, void* cuda_stream) {
    if (!x || !y || !G || !workspace) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_gram: null pointer");
    if (n < 1 || F < 1 || D < 0 || D > MAX_D) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_gram: need n >= 1, F >= 1, 0 <= D <= 16");
    int64_t need = 0;
    int S = 1;
    int rc = qkan_cheb_gram_workspace(n, F, D, &need, &S);
    if (rc) return rc;
    if (workspace_bytes < need) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_gram: workspace too small (see qkan_cheb_gram_workspace)");
    cudaStream_t stream = (cudaStream_t)cuda_stream;
    const GramLayout g = gram_layout(F, D);
    GramParams p;
    p.x = x; p.y = y; p.partial = (double*)workspace; p.n = n; p.F = F; p.D = D; p.DC = g.DC; p.P = g.P; p.W = g.W;
    p.T = g.T;
    p.S = S;
    double* Gw = p.partial + (size_t)S * g.n_tiles * TILE * TILE;
    const dim3 grid(g.n_tiles, S);
    switch (D) {                                             // slots per side <= 64 / DC + 2 features + the y slot
        case 0: qkan_cheb_gram_kernel<17, 1, 0><<<grid, GRAM_THREADS, 0, stream>>>(p); break;
        case 1: qkan_cheb_gram_kernel<17, 1, 1><<<grid, GRAM_THREADS, 0, stream>>>(p); break;
        case 2: qkan_cheb_gram_kernel<9, 2, 1><<<grid, GRAM_THREADS, 0, stream>>>(p); break;
        case 3:
            if (gram_minb(D) == 3) qkan_cheb_gram_kernel<6, 3, 1, 3><<<grid, GRAM_THREADS, 0, stream>>>(p);
            else qkan_cheb_gram_kernel<6, 3, 1, 4><<<grid, GRAM_THREADS, 0, stream>>>(p);
            break;
        case 4:
            if (gram_minb(D) == 3) qkan_cheb_gram_kernel<5, 4, 1, 3><<<grid, GRAM_THREADS, 0, stream>>>(p);
            else qkan_cheb_gram_kernel<5, 4, 1, 4><<<grid, GRAM_THREADS, 0, stream>>>(p);
            break;
        default:
            if (gram_minb(D) == 3) qkan_cheb_gram_kernel<5, 0, 1, 3><<<grid, GRAM_THREADS, 0, stream>>>(p);
            else qkan_cheb_gram_kernel<5, 0, 1, 4><<<grid, GRAM_THREADS, 0, stream>>>(p);
            break;
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_gram_kernel launch");
    qkan_cheb_gram_reduce_kernel<<<dim3(g.n_tiles, TILE * TILE / 256), 256, 0, stream>>>(p.partial, g.n_tiles, S, g.T, g.W, Gw);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_gram_reduce_kernel launch");
    const long long full = (long long)(F * (D + 1) + 1) * (F * (D + 1) + 1);
    qkan_cheb_gram_expand_kernel<<<(unsigned)((full + 255) / 256), 256, 0, stream>>>(Gw, g.W, F, D, g.DC, g.K0, G);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_gram_expand_kernel launch");
    return QKAN_OK;
}

extern "C" int qkan_cheb_residuals_ctas(int* ctas) {
    if (!ctas) return fail(QKAN_ERR_BAD_SHAPE, "null argument");
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    *ctas = 4 * sms;
    return QKAN_OK;
}

extern "C" int qkan_cheb_residuals(const double* x, const double* y, const double* w, int64_t n, int F, int D, const double* coef,
                                   double ybar, double* sums, double* tail, double* xtr, void* cuda_stream) {
    if (!x || !y || !coef || !sums || !tail) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_residuals: null pointer");
    if (n < 1 || F < 1 || D < 0 || D > MAX_D) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_residuals: need n >= 1, F >= 1, 0 <= D <= 16");
    int ctas = 0;
    qkan_cheb_residuals_ctas(&ctas);
    const int D1 = D + 1, P = F * D1;
    cudaError_t e = cudaSuccess;
    // tile kernel: D <= 4, F <= 128, the lane's X^T r accumulators within 64 doubles, tiles + tables within 200 KiB
    {
        const int jf = (F + 31) / 32;
        // two rounds of 64 samples per tile (each coefficient load serves two samples per lane) when that fits shared memory
        const size_t tsmem2 = (size_t)rt_smem(F, D1, xtr != nullptr, 2 * RT_TS1).total * sizeof(double);
        int spl = tsmem2 <= 200 * 1024 ? 2 : 1;
        if (const char* e = getenv("QKAN_RES_SPL")) spl = (atoi(e) == 2 && tsmem2 <= 200 * 1024) ? 2 : 1;   // A/B aid
        const size_t tsmem = spl == 2 ? tsmem2 : (size_t)rt_smem(F, D1, xtr != nullptr, RT_TS1).total * sizeof(double);
        const char* force = getenv("QKAN_RES_KERNEL");           // A/B aid: "warp" = the warp-per-sample kernel
        const bool warp_only = force && force[0] == 'w';
        if (!warp_only && D1 <= 5 && jf <= 4 && jf * D1 * D1 <= 64 && tsmem <= 200 * 1024) {
            const int tma_ok = (reinterpret_cast<uintptr_t>(x) & 15) == 0;
#define QK_RT_LAUNCH1(D1V, JFV, SPLV)                                                                                               \
    {                                                                                                                                \
        e = cudaFuncSetAttribute(qkan_cheb_residual_tile_kernel<D1V, JFV, SPLV>, cudaFuncAttributeMaxDynamicSharedMemorySize,        \
                                 (int)tsmem);                                                                                        \
        if (e == cudaSuccess)                                                                                                        \
            qkan_cheb_residual_tile_kernel<D1V, JFV, SPLV><<<ctas, RT_THREADS, tsmem, (cudaStream_t)cuda_stream>>>(                  \
                x, y, w, n, F, coef, ybar, sums, tail, xtr, tma_ok);                                                                 \
    }
#define QK_RT_LAUNCH(D1V, JFV)                                                                                                       \
    {                                                                                                                                \
        if (spl == 2) QK_RT_LAUNCH1(D1V, JFV, 2) else QK_RT_LAUNCH1(D1V, JFV, 1)                                                     \
    }
#define QK_RT_CASE(D1V)                                                                                                              \
    case D1V:                                                                                                                        \
        if (jf == 1) QK_RT_LAUNCH(D1V, 1)                                                                                            \
        else if (jf == 2) QK_RT_LAUNCH(D1V, 2)                                                                                       \
        else if (jf == 3) QK_RT_LAUNCH(D1V, (3 * D1V * D1V <= 64 ? 3 : 1))                                                           \
        else QK_RT_LAUNCH(D1V, (4 * D1V * D1V <= 64 ? 4 : 1))                                                                        \
        break;
            switch (D1) { QK_RT_CASE(1) QK_RT_CASE(2) QK_RT_CASE(3) QK_RT_CASE(4) QK_RT_CASE(5) }
#undef QK_RT_CASE
#undef QK_RT_LAUNCH
#undef QK_RT_LAUNCH1
            if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(qkan_cheb_residual_tile_kernel)");
            e = cudaGetLastError();
            if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_residual_tile_kernel launch");
            return QKAN_OK;
        }
    }
    const size_t smem = ((xtr ? (size_t)D1 * P : 0) + (size_t)(RES_THREADS / 32) * (2 * D1 + 4)) * sizeof(double);
    if (smem > 200 * 1024) return fail(QKAN_ERR_UNSUPPORTED, "qkan_cheb_residuals: (D+1)^2 F too large for the refinement accumulators");
#define QK_RES_LAUNCH(D1V, JFV)                                                                                                      \
    {                                                                                                                                \
        e = cudaFuncSetAttribute(qkan_cheb_residual_kernel<D1V, JFV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);       \
        if (e == cudaSuccess)                                                                                                        \
            qkan_cheb_residual_kernel<D1V, JFV><<<ctas, RES_THREADS, smem, (cudaStream_t)cuda_stream>>>(x, y, w, n, F, coef, ybar,   \
                                                                                                     sums, tail, xtr);             \
    }
    // register accumulators for X^T r when a lane's share (ceil(F / 32) features x (D+1)^2) fits 64 doubles
#define QK_RES_CASE(DD)                                                                                                              \
    case DD:                                                                                                                         \
        if (jf == 1 && 1 * (DD + 1) * (DD + 1) <= 64) QK_RES_LAUNCH(DD + 1, (1 * (DD + 1) * (DD + 1) <= 64 ? 1 : 0))          \
        else if (jf == 2 && 2 * (DD + 1) * (DD + 1) <= 64) QK_RES_LAUNCH(DD + 1, (2 * (DD + 1) * (DD + 1) <= 64 ? 2 : 0))     \
        else if (jf == 3 && 3 * (DD + 1) * (DD + 1) <= 64) QK_RES_LAUNCH(DD + 1, (3 * (DD + 1) * (DD + 1) <= 64 ? 3 : 0))     \
        else if (jf == 4 && 4 * (DD + 1) * (DD + 1) <= 64) QK_RES_LAUNCH(DD + 1, (4 * (DD + 1) * (DD + 1) <= 64 ? 4 : 0))     \
        else QK_RES_LAUNCH(DD + 1, 0)                                                                                                \
        break;
    const int jf = (F + 31) / 32;
    switch (D) {
        QK_RES_CASE(0) QK_RES_CASE(1) QK_RES_CASE(2) QK_RES_CASE(3) QK_RES_CASE(4) QK_RES_CASE(5) QK_RES_CASE(6) QK_RES_CASE(7)
        QK_RES_CASE(8) QK_RES_CASE(9) QK_RES_CASE(10) QK_RES_CASE(11) QK_RES_CASE(12) QK_RES_CASE(13) QK_RES_CASE(14)
        QK_RES_CASE(15) QK_RES_CASE(16)
    }
#undef QK_RES_LAUNCH
#undef QK_RES_CASE
    if (e != cudaSuccess) return cuda_fail(e, "cudaFuncSetAttribute(qkan_cheb_residual_kernel)");
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_residual_kernel launch");
    return QKAN_OK;
}

// ---- FP64 tensor-core peak: independent DMMA accumulator chains on every SM (the Gram kernel's roofline denominator)
namespace {
__global__ void __launch_bounds__(256) dmma_peak_kernel(double* sink, int iters, double a, double b) {
    double c[8][2];
#pragma unroll
    for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = (double)(threadIdx.x + i) * 1e-3;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) dmma_m8n8k4(c[i][0], c[i][1], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
    if (s == 123456.789) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int qkan_measure_dmma_peak(int device, double* tflops) {
    if (!tflops) return fail(QKAN_ERR_BAD_SHAPE, "null argument");
    int prev_device = -1;
    cudaGetDevice(&prev_device);
    struct Restore { int d; ~Restore() { if (d >= 0) cudaSetDevice(d); } } restore{prev_device == device ? -1 : prev_device};
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) return cuda_fail(e, "cudaSetDevice");
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device);
    const int grid = sms * 8, nt = 256, iters = 4096;
    double* sink = nullptr;
    e = cudaMalloc(&sink, (size_t)grid * nt * sizeof(double));
    if (e != cudaSuccess) return cuda_fail(e, "cudaMalloc");
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {
        cudaEventRecord(e0);
        dmma_peak_kernel<<<grid, nt>>>(sink, iters, 1e-3, 1e-3);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0;
        cudaEventElapsedTime(&ms, e0, e1);
        // one m8n8k4 per warp = 8 * 8 * 4 multiply-adds = 512 flops
        const double fl = 512.0 * 8 * 2 * (double)iters * grid * (nt / 32);
        const double tf = fl / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "dmma_peak_kernel");
    *tflops = best;
    return QKAN_OK;
}

extern "C" int qkan_cheb_features(const double* x, int64_t n, int F, int D, double* out, void* cuda_stream) {
    if (!x || !out) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_features: null pointer");
    if (n < 0 || F < 1 || D < 0) return fail(QKAN_ERR_BAD_SHAPE, "qkan_cheb_features: bad shape");
    if (n == 0) return QKAN_OK;
    const long long tot = (long long)n * F;
    qkan_cheb_features_kernel<<<(unsigned)((tot + 255) / 256), 256, 0, (cudaStream_t)cuda_stream>>>(x, n, F, D, out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e, "qkan_cheb_features_kernel launch");
    return QKAN_OK;
}
