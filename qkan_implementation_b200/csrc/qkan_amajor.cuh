// qkan_amajor.cuh - the default forward kernels of the block engine (compat mode, 1 <= D <= 16).
//
// Same circuit, same per-block arithmetic as qkan_block.cuh (scaled rotations, evolve_blocks_tan), but the
// lane's blocks are walked "a-major": the D + 1 blocks (a, b, d = 0 .. D) of one (a, b) are consecutive.
// Their CHEB rotation is the same multiplexor entry - x[(a + N b) / K] does not depend on d
// (ChebyshevStep.py:64, MulStep.py:59) - so the (t, alpha, beta) triple is fetched ONCE per (a, b) and kept
// in registers while the D + 1 blocks are evolved; per block the kernel only streams the SELECT rotation
// (cos, sin)(theta_w / 2) through one pointer with immediate offsets.  Against the slot-major walk of
// round 1 (three LDS.64 + one table-offset load + pointer bumps per block) this removes about
// 3 of 4 non-arithmetic instructions of a shallow sequence (ncu: profiles/r02_ncu_*.txt).
//
// Mapping: one sample = G = G_k * G_r lanes of one warp.  Lane (k, r) owns output rows b = bi G_k + k and,
// in row b, the summed indices a = pi G_r + r, pi = 0 .. passes - 1 (passes = ceil(N / G_r)); all D + 1
// degree blocks of (a, b) are evolved by that lane.  UNPREPARE + SUM + post-selection on deg = a = 0 is the
// lane's running sum, finished across the G_r lanes of the row by an xor butterfly.
//
// Tables (built once per set_weights by qkan_prepare_amajor_tables_kernel):
//     step = ((bi * passes + pi) << g_log2) + g,   g = (k << g_r_log2) | r
//     wtab[step * (D + 1) + d] = (cos, sin)(theta_w / 2) of block (a, b, d): W[d][a + N b] (column-major SUM
//                                reshape, QKANLayer.py:132; MulStep.py:69); a lane's D + 1 entries are contiguous
//     xotab[step]              = byte offset of the (t, alpha, beta) triple of x[(a + N b) / K] in the sample's
//                                cs row (window kernel: relative to the row step's input window)
// Padding steps (a >= N or b >= K, and one extra pass at the end for the prefetch) rotate by theta = pi and read
// the row's dummy entry (x = 0): they add exactly 0.
#pragma once
#include "qkan_block.cuh"

namespace qkan {

// pick (G_r, G_k) for the a-major walk: maximise live / issued (a, b) steps, then few lanes per row (no shuffle
// steps, more samples per warp), then many rows in parallel.
inline BlockLayout plan_amajor_layout(int N, int K, int min_g_log2 = 0, int max_gk_log2 = 5) {
    BlockLayout best{};
    double best_score = -1.0;
    for (int gr = 0; gr <= 5; ++gr) {
        for (int gk = 0; gr + gk <= 5 && gk <= max_gk_log2; ++gk) {
            if (gr + gk < min_g_log2) continue;
            const long long G_r = 1ll << gr, G_k = 1ll << gk;
            const long long passes = (N + G_r - 1) / G_r;
            const long long brows = (K + G_k - 1) / G_k;
            const double eff = (double)((long long)N * K) / (double)(G_r * passes * G_k * brows);
            const double score = eff - 1e-4 * gr + 1e-5 * gk;
            if (score > best_score) {
                best_score = score;
                best.U = 1; best.g_r_log2 = gr; best.g_k_log2 = gk;
                best.passes = (int)passes; best.brows = (int)brows; best.efficiency = eff;
            }
        }
    }
    return best;
}

// table entries of one (bi, pi, lane) step: D + 1 SELECT rotations and the offset of the CHEB triple
template <typename R>
QK_HD void fill_amajor_step(long long step, const double* W, int N, int K, int D, int passes, int brows, int g_r_log2,
                            int g_k_log2, CS<R>* wtab, int* xotab, int x_entry_bytes, int window) {
    const int g_log2 = g_r_log2 + g_k_log2;
    const int g = (int)(step & ((1ll << g_log2) - 1));
    const long long t = step >> g_log2;
    const int pi = (int)(t % passes);
    const int bi = (int)(t / passes);
    const int k = g >> g_r_log2, r = g & ((1 << g_r_log2) - 1);
    const long long b = ((long long)bi << g_k_log2) + k;
    const long long a = ((long long)pi << g_r_log2) + r;
    const bool live = bi < brows && b < K && a < N;
    int xo = (window ? window : N) * x_entry_bytes;           // the row's dummy entry (x = 0)
    long long flat = 0;
    if (live) {
        flat = a + (long long)N * b;
        int wlo = 0, wlen = 0;
        if (window) block_window(N, K, g_k_log2, bi, &wlo, &wlen);
        xo = (int)(flat / K - wlo) * x_entry_bytes;
    }
    xotab[step] = xo;
    for (int d = 0; d <= D; ++d) {
        CS<R> q;
        q.c = R(0); q.s = R(1);
        if (live) {
            const R w = (R)W[(long long)d * N * K + flat];
            q.c = w;
            q.s = qk_sqrt((R(1) - w) * (R(1) + w));
        }
        wtab[step * (D + 1) + d] = q;
    }
}

// steps of the tables: every (row step, pass, lane) plus one pass of padding (the kernels prefetch one pass ahead)
inline long long amajor_steps(const BlockLayout& lay) {
    return ((long long)lay.brows * lay.passes + 1) << (lay.g_r_log2 + lay.g_k_log2);
}

// shared memory of the main kernel for a tile of `sub` sub-iterations (SPC samples each): two raw-x TMA buffers
// (unless the pre-pass reads global memory directly), the cs tile with SU - 1 sub-iterations of slack rows, two
// mbarriers.  Used by the kernel selection AND the launch, so that a layout accepted at create time launches.
inline bool amajor_direct_x(int SPC, int N) { return (size_t)SPC * N * 16 > 16 * 1024; }
inline size_t amajor_smem_bytes(int N, int SPC, int row_bytes, int SU, int sub) {
    const size_t tile = (size_t)SPC * sub;
    const size_t xs = amajor_direct_x(SPC, N) ? 0 : 2 * ((tile * N + 1) & ~(size_t)1) * sizeof(double);
    const size_t cs = ((tile + (size_t)(SU - 1) * SPC) * (size_t)row_bytes + 15) & ~(size_t)15;
    return xs + cs + 16;
}
constexpr size_t AMAJOR_SMEM_CAP = 200 * 1024;

// The D + 1 blocks (a, b, d = 0 .. D) of SU samples: CHEB triples in registers, SELECT rotations streamed from `wp`.
//
// CHEB acts on f_x only and PREPARE leaves deg in the product state |+>^l, so until SELECT the statevector is
// (block state of (a, b)) (x) |+>_deg: the D + 1 blocks of one (a, b) hold the SAME four amplitudes.  With the triple
// in registers their D + 1 evolutions are literally common subexpressions (ptxas merges them whether or not the source
// spells it out - an earlier version that evolved every block separately compiled to this code), so the function says
// it explicitly: the block state is evolved ONCE through the CHEB sequence (8 (D - 1) FMA + the pruned last pass,
// 4 MUL + 4 FMA) and the SELECT rotation is then applied to each of the D + 1 copies with its own angle (4 FMA per
// block, fused with the read-out sum).  Exact; no weight is pre-summed - every (a, b, d) amplitude gets its own SELECT
// rotation.  Per (a, b): 12 D + 4 FP instructions / 24 D + 4 flops (complex amplitudes), against (8 D + 4)(D + 1)
// when every block is evolved on its own (round 1).
template <class A, typename R, int SU, int DT>
QK_HD void amajor_blocks(const A (&init)[4], const TanEntry<R> (&e)[SU], const CS<R>* __restrict__ wp,
                                              A (&acc)[SU]) {
    A lo0[SU], lo2[SU];
    QK_UNROLL
    for (int j = 0; j < SU; ++j) {
        A v[4];
        QK_UNROLL
        for (int q = 0; q < 4; ++q) v[q] = init[q];
        QK_UNROLL
        for (int r = 0; r + 1 < DT; ++r) {
            rot_tan(v[0], v[1], e[j].t);
            rot_tan(v[2], v[3], e[j].t);
        }
        lo0[j] = lin2(v[0], v[1], e[j].al, e[j].be);
        lo2[j] = lin2(v[2], v[3], e[j].al, e[j].be);
    }
    QK_UNROLL
    for (int d = 0; d <= DT; ++d) {
        const CS<R> q = wp[d];
        QK_UNROLL
        for (int j = 0; j < SU; ++j) {
            fma_amp(acc[j], q.c, lo0[j]);
            fma_amp(acc[j], -q.s, lo2[j]);
        }
    }
}

#if defined(__CUDACC__)
template <class A, typename R>
__device__ __forceinline__ void amajor_store(const BlockParams& p, const A& acc, long long o, long long oa) {
    store_result(p, o, (double)acc.re * p.out_scale);
    if (p.amps) {
        Cplx<R> z;
        z.re = (R)((double)acc.re * p.amp_scale);
        if constexpr (A::is_complex) z.im = (R)((double)acc.im * p.amp_scale);
        else z.im = R(0);
        reinterpret_cast<Cplx<R>*>(p.amps)[oa] = z;
    }
}

// SIMPLE: every lane owns one whole output row of its sample (G_r = 1 and a single row step - the layout of
// every small BASELINE layer): no row loop and no shuffle step in the per-sample code.
template <class A, typename R, int SU, int NT, int MINB, bool SIMPLE, int DT>
__global__ void __launch_bounds__(NT, MINB) qkan_block_amajor_kernel(const BlockParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = p.G, G_r = p.G_r, G_k = p.G_k;
    const int SPC = p.SPC;                                   // samples in flight per CTA
    const int tile = p.tile;                                 // samples per x tile
    const int RB = p.row_bytes;                              // cs row stride: N triples + the dummy (+ padding)
    constexpr size_t ENTB = sizeof(TanEntry<R>);
    constexpr int D1 = DT + 1;
    // smem: xs[2] (TMA destinations: raw x rows, two tiles in flight) | cs (rotation triples of the current tile, plus
    // SU - 1 sub-iterations of slack rows: the idle slots of a ragged tile read past its last row) | mbar[2]
    const size_t xs_doubles = p.direct_x ? 0 : (((size_t)tile * p.N + 1) & ~(size_t)1);
    double* xs0 = reinterpret_cast<double*>(smem_raw);
    char* cs = reinterpret_cast<char*>(smem_raw + 2 * xs_doubles * sizeof(double));
    unsigned long long* mbar =
        reinterpret_cast<unsigned long long*>(smem_raw + 2 * xs_doubles * sizeof(double) + (((size_t)(tile + (SU - 1) * SPC) * RB + 15) & ~(size_t)15));

    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int r = g & (G_r - 1);
    const int k = g >> p.g_r_log2;
    const int slot = tid >> (p.g_r_log2 + p.g_k_log2);       // sample slot inside the CTA
    // this CTA's slice of the batch: samples [base, bend), walked tile by tile (s_tot < 0: tiles of the whole batch
    // dealt round-robin to the CTAs)
    const bool strided = p.s_tot < 0;
    const long long base = strided ? 0 : ((long long)blockIdx.x * p.s_tot / gridDim.x) * SPC;
    const long long bend_raw = strided ? p.B : (((long long)blockIdx.x + 1) * p.s_tot / gridDim.x) * SPC;
    const long long bend = bend_raw < p.B ? bend_raw : p.B;
    const long long it_step = strided ? (long long)gridDim.x : 1;
    const long long n_it = (bend - base + tile - 1) / tile;
    const CS<R>* __restrict__ wtab = reinterpret_cast<const CS<R>*>(p.cstab);
    const int* __restrict__ xotab = p.xotab;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    for (int i = tid; i < tile + (SU - 1) * SPC; i += NT) {   // the dummy entries never change; slack rows are all-dummy
        const TanEntry<R> dm = tan_entry<R>(R(0), DT);
        if (i < tile) {
            *reinterpret_cast<TanEntry<R>*>(cs + (size_t)i * RB + p.N * ENTB) = dm;
        } else {
            for (int n = 0; n <= p.N; ++n) *reinterpret_cast<TanEntry<R>*>(cs + (size_t)i * RB + n * ENTB) = dm;
        }
    }
    __syncthreads();

    auto tile_bytes = [&](long long it) -> unsigned {
        const long long s0 = base + it * tile;
        const int ns = (int)((bend - s0 < tile) ? (bend - s0) : tile);
        return (unsigned)ns * (unsigned)p.N * 8u;
    };
    // stage the x rows of tile `it` into buffer b: one 1-D TMA bulk copy when the tile is 16-byte
    // granular, plain coalesced loads otherwise (ragged tail, odd N)
    auto issue_x = [&](long long it, int b) {
        if (p.direct_x) return;
        const unsigned bytes = tile_bytes(it);
        const double* src = p.x + (base + it * tile) * p.N;
        double* dst = xs0 + (size_t)b * xs_doubles;
        if (p.tma_ok && (bytes & 15u) == 0) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&mbar[b], bytes);
                tma_load_1d(dst, src, bytes, &mbar[b]);
            }
        } else {
            for (int i = tid; i < (int)(bytes >> 3); i += NT) dst[i] = src[i];
        }
    };

    A init[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        init[q].re = (R)p.init[2 * q];
        if constexpr (A::is_complex) init[q].im = (R)p.init[2 * q + 1];
    }
    const int x0 = xotab[g];                                  // first step's offset: resident for the whole launch
    const size_t wstep = (size_t)G * D1;                      // table entries between consecutive passes of a lane

    // two tiles in flight: tile i is consumed while tiles i+1 and (after its pre-pass) i+2 are loading
    long long it = strided ? (long long)blockIdx.x : 0;
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    if (it < n_it) issue_x(it, 0);
    if (it + it_step < n_it) issue_x(it + it_step, 1);
    __syncthreads();
    // pre-pass walk: thread tid takes inputs tid, tid + NT, ... of the tile; its (row, n) advances by
    // (NT / N, NT % N) per step, so the loop has no division
    const int pre_row0 = tid / p.N, pre_n0 = tid - pre_row0 * p.N;
    const int pre_dr = NT / p.N, pre_dn = NT - pre_dr * p.N;

    const size_t row_stride = (size_t)SPC * RB;                   // bytes between consecutive sub-iterations
    const long long out_stride = (long long)SPC * p.K;

    for (; it < n_it; it += it_step, buf ^= 1) {
        if (!p.direct_x && p.tma_ok && (tile_bytes(it) & 15u) == 0) {
            if (buf == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
        }
        const long long s0 = base + it * tile;
        const double* xs = p.direct_x ? p.x + s0 * p.N : xs0 + (size_t)buf * xs_doubles;
        const int nsamp = (int)((bend - s0 < tile) ? (bend - s0) : tile);

        // pre-pass over the raw inputs of the tile: range count (the reference prints a warning,
        // ChebyshevStep.py:46-49), clip (:52) and the scaled-rotation triple of cos(theta/2) = x - no arccos
        unsigned bad = 0;
        {
            const int n_in = nsamp * p.N;
            int row = pre_row0, n = pre_n0;
            for (int e = tid; e < n_in; e += NT) {
                const double v = xs[e];
                if (!(-1.0 - 1e-8 <= v) || !(v <= 1.0 + 1e-8)) ++bad;
                *reinterpret_cast<TanEntry<R>*>(cs + (size_t)row * RB + n * ENTB) = tan_entry<R>(clip_unit<R>(v), DT);
                n += pre_dn;
                row += pre_dr;
                if (n >= p.N) { n -= p.N; ++row; }
            }
        }
        if (bad) atomicAdd(p.oor, (unsigned long long)bad);
        __syncthreads();                                      // cs complete, xs[buf] free again
        const long long nxt = it + 2 * it_step;
        if (nxt < n_it) issue_x(nxt, buf);                    // overlaps with the compute of this and the next tile

        // SU samples per lane at a time (the lane's slots of SU consecutive sub-iterations): they share every
        // SELECT entry and the per-pass bookkeeping
        const int nsub = (nsamp + SPC - 1) / SPC;
        const char* csrow = cs + (size_t)slot * RB;
        long long o = (p.row0 + s0 + slot) * p.K;
        long long oa = (s0 + slot) * p.K;                     // amps are local: no row offset
        int ls = slot;
        for (int si = 0; si < nsub; si += SU, csrow += SU * row_stride, o += SU * out_stride, oa += SU * out_stride, ls += SU * SPC) {
            bool valid[SU];
            const char* row[SU];
            QK_UNROLL
            for (int j = 0; j < SU; ++j) {
                valid[j] = ls + j * SPC < nsamp;
                row[j] = csrow + j * row_stride;      // idle slots of a ragged tile evolve a stale / slack row; nothing is stored
            }
            const CS<R>* wp = wtab + (size_t)g * D1;
            const int* xp = xotab + g;
            int xo = x0;
            A acc[SU];
            auto run_row = [&]() {
                QK_UNROLL
                for (int j = 0; j < SU; ++j) set_amp(acc[j], 0.0);
                for (int pi = 0; pi < p.passes; ++pi) {
                    TanEntry<R> e[SU];
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) e[j] = *reinterpret_cast<const TanEntry<R>*>(row[j] + xo);
                    xp += G;
                    xo = *xp;                                 // next pass (the tables end with one pass of padding steps)
                    amajor_blocks<A, R, SU, DT>(init, e, wp, acc);
                    wp += wstep;
                }
            };
            if constexpr (SIMPLE) {
                run_row();
                QK_UNROLL
                for (int j = 0; j < SU; ++j)
                    if (valid[j] && k < p.K) amajor_store<A, R>(p, acc[j], o + j * out_stride + k, oa + j * out_stride + k);
            } else {
                for (int b = k; b < p.brows * G_k; b += G_k) {
                    run_row();
                    // UNPREPARE (H on deg) + SUM (H on a) + post-selection deg = a = 0: the sum over the
                    // row's blocks, finished across the G_r lanes with an xor butterfly
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) {
                        for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc[j], shfl_xor_amp(acc[j], m));
                        if (valid[j] && r == 0 && b < p.K) amajor_store<A, R>(p, acc[j], o + j * out_stride + b, oa + j * out_stride + b);
                    }
                }
            }
        }
        __syncthreads();                                      // everyone done with cs before the next pre-pass
    }
}

// Window kernel: wide input rows (N784 K10: 6.3 KB of x, 19 KB of rotation triples per sample).  The triples of
// a sample are built per ROW STEP from the step's input window (block_window) instead of once per sample, so a
// CTA keeps tile * (W + 1) triples instead of tile * (N + 1) and shared memory no longer limits the resident
// warps.  x is read straight from global memory (each input once per row step that uses it: twice at most, at
// window boundaries).  One sample per lane at a time.
template <class A, typename R, int NT, int MINB, int DT>
__global__ void __launch_bounds__(NT, MINB) qkan_block_amajor_window_kernel(const BlockParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = p.G, G_r = p.G_r;
    const int SPC = p.SPC, tile = p.tile, RB = p.row_bytes, W = p.window;
    constexpr size_t ENTB = sizeof(TanEntry<R>);
    constexpr int D1 = DT + 1;
    char* cs = reinterpret_cast<char*>(smem_raw);
    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int r = g & (G_r - 1);
    const int k = g >> p.g_r_log2;
    const int slot = tid >> (p.g_r_log2 + p.g_k_log2);
    const CS<R>* __restrict__ wtab = reinterpret_cast<const CS<R>*>(p.cstab);
    const int* __restrict__ xotab = p.xotab;

    for (int i = tid; i < tile; i += NT)                      // the dummy entries never change
        *reinterpret_cast<TanEntry<R>*>(cs + (size_t)i * RB + W * ENTB) = tan_entry<R>(R(0), DT);
    A init[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        init[q].re = (R)p.init[2 * q];
        if constexpr (A::is_complex) init[q].im = (R)p.init[2 * q + 1];
    }
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        keep_in_register(init[q].re);
        if constexpr (A::is_complex) keep_in_register(init[q].im);
    }
    const long long n_it = (p.B + tile - 1) / tile;
    const size_t step_steps = (size_t)p.passes * G;           // table steps of one row step
    const size_t wstep = (size_t)G * D1;

    for (long long it = blockIdx.x; it < n_it; it += gridDim.x) {
        const long long s0 = it * tile;
        const int nsamp = (int)((p.B - s0 < tile) ? (p.B - s0) : tile);
        const int nsub = (nsamp + SPC - 1) / SPC;
        int prev_hi = -1;
        for (int bi = 0; bi < p.brows; ++bi) {
            int lo, len;
            block_window(p.N, p.K, p.g_k_log2, bi, &lo, &len);
            __syncthreads();                                  // the previous row step's entries are consumed
            // pre-pass over the window of every sample of the tile (flat walk, no division in the loop): range count
            // (ChebyshevStep.py:46-49; an input shared by two windows is counted once), clip (:52), triple
            unsigned bad = 0;
            {
                const int n_in = nsamp * len;
                int row = tid / len, j = tid - row * len;
                const int dr = NT / len, dj = NT - dr * len;
                const double* xw = p.x + s0 * p.N + lo;
                for (int e = tid; e < n_in; e += NT) {
                    const double v = xw[(size_t)row * p.N + j];
                    if (lo + j > prev_hi && (!(-1.0 - 1e-8 <= v) || !(v <= 1.0 + 1e-8))) ++bad;
                    *reinterpret_cast<TanEntry<R>*>(cs + (size_t)row * RB + j * ENTB) = tan_entry<R>(clip_unit<R>(v), DT);
                    j += dj;
                    row += dr;
                    if (j >= len) { j -= len; ++row; }
                }
            }
            prev_hi = lo + len - 1;
            if (bad) atomicAdd(p.oor, (unsigned long long)bad);
            __syncthreads();

            const int b = (bi << p.g_k_log2) + k;
            const CS<R>* wp0 = wtab + ((size_t)bi * step_steps + g) * D1;
            const int* xp0 = xotab + (size_t)bi * step_steps + g;
            const int x0 = xp0[0];
            int ls = slot;
            for (int si = 0; si < nsub; ++si, ls += SPC) {
                const char* row = cs + (size_t)ls * RB;       // idle slots of a ragged tile evolve a stale row; nothing is stored
                const CS<R>* wp = wp0;
                const int* xp = xp0;
                int xo = x0;
                A acc[1];
                set_amp(acc[0], 0.0);
                for (int pi = 0; pi < p.passes; ++pi) {
                    TanEntry<R> e[1];
                    e[0] = *reinterpret_cast<const TanEntry<R>*>(row + xo);
                    xp += G;
                    xo = *xp;                                 // next pass (the tables end with one pass of padding steps)
                    amajor_blocks<A, R, 1, DT>(init, e, wp, acc);
                    wp += wstep;
                }
                // UNPREPARE + SUM + post-selection: the sum over the row's blocks, finished across the G_r lanes
                for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc[0], shfl_xor_amp(acc[0], m));
                if (ls < nsamp && r == 0 && b < p.K)
                    amajor_store<A, R>(p, acc[0], (p.row0 + s0 + ls) * p.K + b, (s0 + ls) * p.K + b);
            }
        }
    }
}

template <typename R>
__global__ void qkan_prepare_amajor_tables_kernel(const double* W, int N, int K, int D, int passes, int brows, int g_r_log2,
                                                  int g_k_log2, int x_entry_bytes, int window, long long steps_total,
                                                  CS<R>* wtab, int* xotab, unsigned long long* bad_weights) {
    const long long step = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    // |w| <= 1 is required for the rotation to exist (MulStep.py:36-37): check each weight once
    const long long nw = (long long)N * K * (D + 1);
    unsigned bad = 0;
    for (long long i = step; i < nw; i += (long long)gridDim.x * blockDim.x)
        if (!(fabs(W[i]) <= 1.0)) ++bad;
    if (bad) atomicAdd(bad_weights, (unsigned long long)bad);
    if (step >= steps_total) return;
    fill_amajor_step<R>(step, W, N, K, D, passes, brows, g_r_log2, g_k_log2, wtab, xotab, x_entry_bytes, window);
}

template <class A, typename R, int SU, int NT, int MINB, bool SIMPLE, int DT>
cudaError_t launch_amajor_impl(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    auto kern = qkan_block_amajor_kernel<A, R, SU, NT, MINB, SIMPLE, DT>;
    BlockParams p = p0;
    const int SPC = NT / G;
    p.row_bytes = tan_row_words(p.N, G, (int)sizeof(R)) * (int)sizeof(R);
    if (const char* e = getenv("QKAN_BLOCK_ROW_WORDS")) {      // tuning aid
        if (atoi(e) >= 3 * (p.N + 1)) p.row_bytes = atoi(e) * (int)sizeof(R);
    }
    // wide rows: staging the raw x twice more than doubles the shared memory per sample and would
    // halve the resident warps; the per-tile compute is long, so the pre-pass reads global memory directly
    p.direct_x = amajor_direct_x(SPC, p.N) ? 1 : 0;
    auto smem_for = [&](int sub) { return amajor_smem_bytes(p.N, SPC, p.row_bytes, SU, sub); };
    int sub = (int)(8192 / ((size_t)SPC * p.N * 8));          // about 8 KiB of x per tile ...
    const int sub_cs = (int)(49152 / ((size_t)SPC * p.row_bytes));   // ... and at most 48 KiB of triples
    if (sub > sub_cs) sub = sub_cs;
    if (const char* e = getenv("QKAN_BLOCK_SUB")) sub = atoi(e);   // tuning aid
    if (sub > 32) sub = 32;
    // a lane takes SU sub-iterations at a time: a tile of an odd number of them would leave a sample slot idle
    auto round_su = [](int v) { v -= v % SU; return v < SU ? SU : v; };
    sub = round_su(sub);
    if (smem_for(sub) > AMAJOR_SMEM_CAP) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(sub));
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_for(sub));
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long resident = (long long)sm_count * per_sm;
    // every CTA owns an equal, contiguous run of sub-iterations (SPC samples each); tiles of `sub` sub-iterations
    // inside it, at least four per CTA so that the x tiles pipeline
    const long long s_tot = (p.B + SPC - 1) / SPC;
    long long grid = resident < s_tot ? resident : s_tot;
    if (grid < 1) grid = 1;
    const long long spc = s_tot / grid;                       // sub-iterations per CTA (some get one more)
    while (sub > SU && spc < 4ll * sub) sub = round_su(sub >> 1);
    p.s_tot = s_tot;
    // few sub-iterations per CTA: one more or less is a visible imbalance between SMs, and dealing the tiles
    // round-robin spreads the remainder over the SMs (measured on N8 K8 D16 with 100 k samples: +4.5 %)
    bool strided = spc < 32;
    if (const char* e2 = getenv("QKAN_BLOCK_STRIDED")) strided = atoi(e2) != 0;   // A/B aid
    if (strided) {
        const long long n_it = (p.B + (long long)SPC * sub - 1) / ((long long)SPC * sub);
        grid = resident < n_it ? resident : n_it;
        p.s_tot = -1;
    }
    p.sub = sub;
    p.tma_ok = ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0 && (((size_t)SPC * p.N * 8) & 15u) == 0) ? 1 : 0;
    p.G = G; p.G_r = 1 << p.g_r_log2; p.G_k = 1 << p.g_k_log2;
    p.SPC = SPC; p.tile = SPC * sub;
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = (int)smem_for(sub);
    kern<<<(unsigned)grid, NT, smem_for(sub), stream>>>(p);
    return cudaGetLastError();
}
template <class A, typename R, int SU, int NT, int MINB, int DT>
cudaError_t launch_amajor(const BlockParams& p, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (p.D != DT) return cudaErrorInvalidValue;
    if (p.g_r_log2 == 0 && p.brows == 1)
        return launch_amajor_impl<A, R, SU, NT, MINB, true, DT>(p, G, sm_count, stream, grid_out, smem_out);
    return launch_amajor_impl<A, R, SU, NT, MINB, false, DT>(p, G, sm_count, stream, grid_out, smem_out);
}
template <class A, typename R, int SU, int NT, int MINB, int DT>
BlockKernelInfo make_amajor_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = 0; k.U = 1; k.SU = SU; k.NT = NT; k.MINB = MINB; k.DT = DT; k.is_default = is_default;
    k.tan = 1; k.window = 0; k.amajor = 1;
    k.launch = &launch_amajor<A, R, SU, NT, MINB, DT>;
    return k;
}

inline size_t amajor_window_smem_bytes(int SPC, int row_bytes, int sub) { return (size_t)SPC * sub * row_bytes; }

template <class A, typename R, int NT, int MINB, int DT>
cudaError_t launch_amajor_window(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (p0.D != DT || p0.window < 1) return cudaErrorInvalidValue;
    auto kern = qkan_block_amajor_window_kernel<A, R, NT, MINB, DT>;
    BlockParams p = p0;
    const int SPC = NT / G;
    p.row_bytes = tan_row_words(p.window, G, (int)sizeof(R)) * (int)sizeof(R);
    int sub = (int)(32768 / ((size_t)SPC * p.row_bytes));      // about 32 KiB of rotation triples per CTA
    if (const char* e = getenv("QKAN_BLOCK_SUB")) sub = atoi(e);   // tuning aid
    if (sub > 32) sub = 32;
    if (sub < 1) sub = 1;
    auto smem_for = [&](int sb) { return amajor_window_smem_bytes(SPC, p.row_bytes, sb); };
    if (smem_for(sub) > AMAJOR_SMEM_CAP) return cudaErrorInvalidConfiguration;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_for(sub));
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem_for(sub));
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long resident = (long long)sm_count * per_sm;
    while (sub > 1 && (p.B + (long long)SPC * sub - 1) / ((long long)SPC * sub) < 4 * resident) sub >>= 1;
    const long long n_it = (p.B + (long long)SPC * sub - 1) / ((long long)SPC * sub);
    long long grid = resident < n_it ? resident : n_it;
    if (grid < 1) grid = 1;
    p.sub = sub;
    p.G = G; p.G_r = 1 << p.g_r_log2; p.G_k = 1 << p.g_k_log2;
    p.SPC = SPC; p.tile = SPC * sub;
    p.tma_ok = 0; p.direct_x = 1; p.s_tot = -1;
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = (int)smem_for(sub);
    kern<<<(unsigned)grid, NT, smem_for(sub), stream>>>(p);
    return cudaGetLastError();
}
template <class A, typename R, int NT, int MINB, int DT>
BlockKernelInfo make_amajor_window_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = 0; k.U = 1; k.SU = 1; k.NT = NT; k.MINB = MINB; k.DT = DT; k.is_default = is_default;
    k.tan = 1; k.window = 1; k.amajor = 1;
    k.launch = &launch_amajor_window<A, R, NT, MINB, DT>;
    return k;
}
#endif  // __CUDACC__

}  // namespace qkan
