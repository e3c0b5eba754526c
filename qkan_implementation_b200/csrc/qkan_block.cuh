// qkan_block.cuh - the "block" engine: the default forward kernel.
//
// The index qubits (a, b, deg) of the QKAN circuit are only ever CONTROLS between the
// PREPARE layer and the UNPREPARE/SUM layer, so the multiplexor gates are block diagonal:
// the 2^q statevector is a direct sum of N*K*(D+1) live blocks of four amplitudes
// (f_x, f_w), one per (b, a, d); blocks of padded index values carry exactly zero into
// the read-out and are not simulated.  Per block the kernel
//   * starts from the PREPARE'd state (1, 0, 0, 0)              (closed form, un-normalised)
//   * applies the D CHEB rotations Ry(theta_x) on f_x           (both f_w halves, complex)
//   * applies the MUL / SELECT rotation Ry(theta_w) on f_w
//   * and feeds amplitude (f_x, f_w) = (0, 0) into the UNPREPARE + SUM + post-selection,
//     which for deg = a = 0 is the plain sum over (a, d): a warp-shuffle xor butterfly.
// Only circuit structure is used (block diagonality, closed-form |+> preparation, pruning of
// the final layer to the post-selected outputs); nothing state dependent is skipped: zero
// imaginary parts and the zero f_w = 1 half are evolved like any other amplitude.
//
// Mapping: one sample = a group of G = G_k * G_r lanes of ONE warp (G <= 32).  Lane (k, r)
// owns output rows b = k, k + G_k, ... and, inside row b, the blocks i = a*(D+1)+d with
// i = (pass*U + u)*G_r + r; U blocks (4U complex amplitudes) are in registers at a time.
// The read-out sum runs over the r lanes by __shfl_xor.  The whole state of a small layer
// (N4 K4 D3: 64 blocks = 256 amplitudes over 16 lanes) is register resident at once.
#pragma once
#include "qkan_core.cuh"
#include <stdlib.h>
#if defined(__CUDACC__)
#include "qkan_kernel.cuh"
#endif

namespace qkan {

// lane layout of one sample, chosen on the host (plan_block_layout)
struct BlockLayout {
    int U;                      // blocks per lane in registers
    int g_r_log2, g_k_log2;
    int passes, brows;
    double efficiency;          // live block slots / issued block slots
};

// cs row stride (entries) of the (cos, sin) pair rows.  One LDS phase serves 128 bytes = P lanes (P = 8 for 16-byte
// pairs, 16 for 8-byte pairs); with G < P lanes per sample a phase spans P / G consecutive sample rows, and the
// lanes of a row read G different entries, so the rows must start G entries apart modulo P: stride = G (mod P).
// Only worth it for narrow rows (N + 1 <= 2 P); wider rows keep N + 1.
inline int cs_row_stride(int N, int G, int pair_bytes) {
    const int P = 128 / pair_bytes;
    int np = N + 1;
    if (G < P && np <= 2 * P)
        while (np % P != G % P) ++np;
    return np;
}

// Row stride (words) of the scaled-rotation rows: a row holds N + 1 (t, alpha, beta) triples.  The lanes of
// one shared-memory phase (P = 128 / word bytes lanes) are P / G sample rows x G consecutive triples and read
// the same member of each: word (s * RSW + 3 n) mod P.  Take the stride RSW in [3 (N + 1), 3 (N + 1) + P) with
// the fewest lanes on one bank (no conflicts for G = 4, 8, 16 in FP64: 20 words for N = 4, 40 for N = 8).
inline int tan_row_words(int N, int G, int word_bytes) {
    const int P = 128 / word_bytes;
    const int need = 3 * (N + 1);
    if (G >= P || N + 1 > 2 * P) return need;
    int best = need, best_worst = 1 << 30;
    for (int rsw = need; rsw < need + P; ++rsw) {
        int cnt[32] = {0};
        int worst = 0;
        for (int lane = 0; lane < P; ++lane) {
            const int w = ((lane / G) * rsw + 3 * (lane % G)) % P;
            if (++cnt[w] > worst) worst = cnt[w];
        }
        if (worst < best_worst) { best_worst = worst; best = rsw; }
    }
    return best;
}

// pick (U, G_r, G_k): maximise slot efficiency, then prefer one block per lane at a time (fewest
// registers -> most warps; measured best or equal on every BASELINE config, profiles/r01_tune_*),
// fewer shuffle steps, more rows in parallel.
// min_g_log2 lets the caller force wide groups (few samples per CTA) for very wide inputs.
// max_gk_log2 caps the rows evolved in parallel (the window kernel wants few: its input window grows with them).
inline BlockLayout plan_block_layout(int N, int K, int D, int min_g_log2 = 0, int force_U = 0, int max_gk_log2 = 5) {
    const long long rowlen = (long long)N * (D + 1);
    BlockLayout best{};
    double best_score = -1.0;
    const int Us[3] = {1, 2, 4};
    for (int ui = 0; ui < 3; ++ui) {
        const int U = Us[ui];
        if (force_U && U != force_U) continue;
        for (int gr = 0; gr <= 5; ++gr) {
            for (int gk = 0; gr + gk <= 5 && gk <= max_gk_log2; ++gk) {
                if (gr + gk < min_g_log2) continue;
                const long long G_r = 1ll << gr, G_k = 1ll << gk;
                const long long passes = (rowlen + G_r * U - 1) / (G_r * U);
                const long long brows = (K + G_k - 1) / G_k;
                const double eff = (double)(rowlen * K) / (double)(G_r * U * passes * G_k * brows);
                // efficiency dominates; the rest only breaks near-ties (within 0.5 %)
                const double score = eff - 1e-3 * (U - 1) - 1e-4 * gr + 1e-5 * gk;
                if (score > best_score) {
                    best_score = score;
                    best.U = U; best.g_r_log2 = gr; best.g_k_log2 = gk;
                    best.passes = (int)passes; best.brows = (int)brows; best.efficiency = eff;
                }
            }
        }
    }
    return best;
}

// Input window of a row step.  Output row b reads the inputs x[(a + N b) / K], a < N: a run of about N / K
// consecutive entries, so row step bi (rows bi G_k .. bi G_k + G_k - 1) needs only the window
// [lo, lo + len) of the sample's input row.  The window kernel builds its rotation entries per row step
// instead of per sample, which divides the shared memory per sample by about K / G_k (N784 K10: 19 KB -> 3.8 KB).
QK_HD void block_window(int N, int K, int g_k_log2, int bi, int* lo, int* len) {
    const long long b0 = (long long)bi << g_k_log2;
    long long b1 = b0 + (1ll << g_k_log2);
    if (b1 > K) b1 = K;
    const long long first = (b0 * N) / K, last = (b1 * N - 1) / K;
    *lo = (int)first;
    *len = (int)(last - first + 1);
}
inline int block_window_max(int N, int K, int g_k_log2, int brows) {
    int w = 1;
    for (int bi = 0; bi < brows; ++bi) {
        int lo, len;
        block_window(N, K, g_k_log2, bi, &lo, &len);
        if (len > w) w = len;
    }
    return w;
}

struct BlockParams {
    const double* x;            // [B, N]
    const void* cstab;          // CS<R>[slots + U*G]: SELECT rotation (cos, sin)(theta_w / 2) per block slot
    const int* xotab;           // int[slots + U*G]: byte offset of the CHEB rotation pair in the sample's cs row
                                //                   (| degree << 24 in paper mode)
    double* outs[8];            // [row0 + B, K] result buffers: outs[0] is local; outs[1..n_out) are the same buffer
                                // of the NVLink peers (fused output gather: every result is stored to every rank)
    int n_out;
    double* mc_out;             // optional NVLink-multicast (NVLS) mapping of the result buffer: one multimem.st per
                                // result is replicated to every rank by the NVSwitch (then outs[] is not written)
    long long row0;             // first row of this rank's slice in the result buffers
    void* amps;                 // optional [B, K] complex (local only)
    unsigned long long* oor;
    long long B;
    int N, K, D;
    int g_r_log2, g_k_log2;     // lanes per row / rows in parallel inside a group
    int passes;                 // ceil(N (D+1) / (G_r * U))
    int brows;                  // ceil(K / G_k)
    int sub;                    // sub-iterations per x tile
    int tma_ok;
    int direct_x;               // wide input rows: no raw-x staging buffers, the pre-pass reads x from global memory
    int window;                 // window kernel: entries per cs row (the widest row-step window, block_window_max); 0 = off
    // derived launch constants (filled by launch_block; kept in the constant bank so that the kernel does
    // not spend registers or instructions on them)
    int G, G_r, G_k;            // lanes per sample / per row / rows in parallel
    int SPC, tile;              // samples in flight per CTA; samples per x tile (SPC * sub)
    int row_bytes;              // cs row stride: N + 1 entries, padded so that the lanes of one shared-memory
                                // phase (128 bytes) hit different banks (see cs_row_stride / amajor_row_amps)
    int plane_bytes;            // a-major kernels: offset of the lo2 plane inside a cs row
    int plain;                  // a-major kernels: results go to outs[0] only (no peers, no multicast, no amplitudes)
    long long s_tot;            // sub-iterations (SPC samples each) of the batch; CTA c of g owns the contiguous run
                                // [c s_tot / g, (c+1) s_tot / g): sizes differ by at most one sub-iteration, so there
                                // is no tail imbalance from whole tiles
    double out_scale, amp_scale;
    double init[8];             // prepared block state, 4 complex amplitudes (re, im): (1,0,0,0) un-normalised
};

// Block-slot tables, laid out in the order the kernel walks them with the lane as the fastest index:
//     slot = (((bi * passes + pi) * U + u) << g_log2) + g,     g = (k << g_r_log2) | r
// so that one warp load reads G consecutive entries (one 128-byte line of (cos, sin) pairs for 8
// lanes) and a lane advances a single pointer by G per block.  Slot (bi, pi, u, k, r) is block
// i = (pi * G_r + r) * U + u of output row b = bi * G_k + k; i < N (D+1) is the block
// (a, d) = (i / (D+1), i % (D+1)): weight W[d][a + N b] (column-major SUM reshape, QKANLayer.py:132;
// MulStep.py:69) and input x[(a + N b) / K] (np.repeat dilation, ChebyshevStep.py:64).  Padding slots
// rotate by theta = pi (c = 0) and read the row's dummy (0, 1) entry: they add exactly 0.
template <typename R>
QK_HD void fill_block_slot(long long slot, const double* W, int N, int K, int D, int U, int passes, int g_r_log2,
                           int g_k_log2, int paper, CS<R>* cstab, int* xotab, int x_entry_bytes = (int)sizeof(CS<R>),
                           int window = 0) {
    const int g_log2 = g_r_log2 + g_k_log2;
    const int g = (int)(slot & ((1ll << g_log2) - 1));
    long long t = slot >> g_log2;
    const int u = (int)(t % U); t /= U;
    const int pi = (int)(t % passes);
    const int bi = (int)(t / passes);
    const int k = g >> g_r_log2, r = g & ((1 << g_r_log2) - 1);
    const int b = (bi << g_k_log2) + k;
    const long long i = ((long long)(pi << g_r_log2) + r) * U + u;
    CS<R> q;
    q.c = R(0); q.s = R(1);
    // dummy entry (x = 0) at the end of every cs row; window mode: rows hold the row step's input window only
    int wlo = 0, wlen = 0;
    if (window) block_window(N, K, g_k_log2, bi, &wlo, &wlen);
    int xo = (window ? window : N) * x_entry_bytes;
    if (b < K && i < (long long)N * (D + 1)) {
        const int a = (int)(i / (D + 1)), d = (int)(i - (long long)a * (D + 1));
        const int flat = a + N * b;
        const R w = (R)W[(long long)d * N * K + flat];
        q.c = w;
        q.s = qk_sqrt((R(1) - w) * (R(1) + w));
        xo = (flat / K - wlo) * x_entry_bytes;
        if (paper) xo |= d << 24;
    }
    cstab[slot] = q;
    xotab[slot] = xo;
}

// first output of a rotation pass only: u' = c u - s v (the f = 0 member of the pair)
template <typename R> QK_HD Cplx<R> rot_lo(const Cplx<R>& u, const Cplx<R>& v, R c, R s) {
    Cplx<R> o;
    o.re = qk_fma(c, u.re, -(s * v.re));
    o.im = qk_fma(c, u.im, -(s * v.im));
    return o;
}
template <typename R> QK_HD Real<R> rot_lo(const Real<R>& u, const Real<R>& v, R c, R s) {
    Real<R> o;
    o.re = qk_fma(c, u.re, -(s * v.re));
    return o;
}

// Evolve U blocks from the prepared state `init` (amplitudes v[fx + 2 fw]; a run-time value: the
// kernel makes no use of it being (1, 0, 0, 0), real, or half zero) and return the sum of their
// post-selected (f_x, f_w) = (0, 0) amplitudes.
//   CHEB  applications 1 .. D-1: full rotation passes on both f_w halves       16 instr / 24 flops
//   CHEB  application D and MUL: pruned to the backward light cone of the post-selected
//         amplitude - (0,0) after MUL needs only the f_x = 0 outputs of the last CHEB pass
//                                                                             8 + 4 instr / 12 + 6 flops
//   read-out: one complex add per block                                        2 instr / 2 flops
// (complex amplitudes; half of that for the real-only representation).
// DT > 0: D is the compile-time constant DT (fully unrolled); DT = 0: run-time D.
template <class A, typename R, int U, int MODE, int DT = 0>
QK_HD A evolve_blocks(const A (&init)[4], const R (&cx)[U], const R (&sx)[U], const R (&cw)[U], const R (&sw)[U],
                      const int (&deg)[U], int D) {
    const int Dv = DT > 0 ? DT : D;
    A v[U][4];
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        QK_UNROLL
        for (int q = 0; q < 4; ++q) v[u][q] = init[q];
    }
    // CHEB: the input block-encoding applied D times, U and Z U^dagger Z alternating; as real
    // matrices both equal Ry(theta_x), so each application is the same rotation pass
    auto coef = [&](int u, int r, R& c, R& s) {
        c = cx[u]; s = sx[u];
        if constexpr (MODE == 1) {                      // paper: term d gets d applications
            const bool on = deg[u] >= r + 1;
            c = on ? c : R(1);
            s = on ? s : R(0);
        }
    };
    auto cheb = [&](int r) {
        QK_UNROLL
        for (int u = 0; u < U; ++u) {
            R c, s;
            coef(u, r, c, s);
            rot(v[u][0], v[u][1], c, s);
            rot(v[u][2], v[u][3], c, s);
        }
    };
    if constexpr (DT > 0) {
        QK_UNROLL
        for (int r = 0; r + 1 < DT; ++r) cheb(r);
    } else {
        for (int r = 0; r + 1 < Dv; ++r) cheb(r);
    }
    A acc;
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        A lo0 = v[u][0], lo2 = v[u][2];
        if (Dv > 0) {                                   // last CHEB application, f_x = 0 outputs only
            R c, s;
            coef(u, Dv - 1, c, s);
            lo0 = rot_lo(v[u][0], v[u][1], c, s);
            lo2 = rot_lo(v[u][2], v[u][3], c, s);
        }
        const A z = rot_lo(lo0, lo2, cw[u], sw[u]);     // MUL / SELECT on f_w, (0,0) output
        if (u == 0) acc = z;
        else add_amp(acc, z);
    }
    return acc;
}

// ---------------------------------------------------------------------------------------------
// Scaled-rotation form of the CHEB sequence (compat mode, compile-time degree DT >= 1).
//
// Ry(theta_x) = [[c, -s], [s, c]] is a scalar times a matrix with a unit diagonal:
//     |c| >= s :  Ry = c * M(t),       t =  s / c,    M(t) = [[1, -t], [t, 1]]
//     |c| <  s :  Ry = s * J * M(t),   t = -c / s,    J = Ry(pi) = [[0, -1], [1, 0]]   (|t| <= 1 either way)
// M(t) costs one FMA per real output (u - t v, v + t u) where the (c, s) form needs a multiply and an
// FMA, so a full CHEB pass over a block is 8 FP64 instructions instead of 16.  Rotations about one
// axis commute, hence Ry^D = gamma^D * J^D * M(t)^D: the scalar gamma^D (gamma = c or s) and the
// quarter turns J^D (a signed swap fixed by D mod 4) are deferred to the last, pruned pass, which
// evaluates the f_x = 0 output of gamma^D J^D M(t) as alpha u + beta v.  This is the same device as
// folding the 1/sqrt(2) of every Hadamard into the read-out scale; every amplitude of every block
// is still evolved through every gate.  |t| <= 1 bounds the un-normalised state by 2^(D/2), which
// is why the form is used for the degree-specialised kernels (D <= 16) only.
constexpr int TAN_MIN_DT = 1;        // D = 1: no full pass to save, but the pruned pass + SELECT are 12 instead of 14 instructions
constexpr int TAN_MAX_DT = 16;       // |t| <= 1 bounds the un-normalised state by 2^(D/2)
// (the kernels that use this form are in qkan_amajor.cuh)

template <typename R> struct TanEntry { R t, al, be; };

#if defined(__CUDA_ARCH__)
QK_HD double qk_rsqrt(double a) { return rsqrt(a); }
QK_HD float qk_rsqrt(float a) { return rsqrtf(a); }
#else
QK_HD double qk_rsqrt(double a) { return 1.0 / sqrt(a); }
QK_HD float qk_rsqrt(float a) { return 1.0f / sqrtf(a); }
#endif

// per input element, once per sample (pre-pass): c = clipped x.  One reciprocal square root serves both
// cases: 1 / s for the quarter-turn case, 1 / (s |c|) otherwise (t = s / c = s^2 / (s c)); s^2 is either 0 or
// >= 2^-53 (2^-24 in FP32), so the tiny bias only matters for s = 0, where it turns 0 * inf into 0.
template <typename R> QK_HD TanEntry<R> tan_entry(R c, int D) {
    const R q = (R(1) - c) * (R(1) + c);                 // s^2
    const R c2 = c * c;
    const bool quarter = c2 < q;                         // |c| < s
    const R tiny = sizeof(R) == 8 ? (R)1e-300 : (R)1e-30;
    const R r = qk_rsqrt(quarter ? q : qk_fma(q, c2, tiny));
    const R gam = quarter ? q * r : c;
    const R t = (quarter ? -c : (c < R(0) ? -q : q)) * r;
    R g = R(1), b = gam;                                 // g = gam^D by squaring (D is a compile-time constant in the kernel)
    for (int e = D; e > 0; e >>= 1) {
        if (e & 1) g *= b;
        if (e > 1) b *= b;
    }
    const R gt = g * t;
    TanEntry<R> e;
    e.t = t;
    // f_x = 0 output of g J^D (u - t v, v + t u)
    switch (quarter ? (D & 3) : 0) {
        case 0: e.al = g; e.be = -gt; break;
        case 1: e.al = -gt; e.be = -g; break;
        case 2: e.al = -g; e.be = gt; break;
        default: e.al = gt; e.be = g; break;
    }
    return e;
}

// (u, v) <- (u - t v, v + t u): one FMA per real output
template <typename R> QK_HD void rot_tan(Cplx<R>& u, Cplx<R>& v, R t) {
    const R ur = u.re, ui = u.im;
    u.re = qk_fma(-t, v.re, ur);
    u.im = qk_fma(-t, v.im, ui);
    v.re = qk_fma(t, ur, v.re);
    v.im = qk_fma(t, ui, v.im);
}
template <typename R> QK_HD void rot_tan(Real<R>& u, Real<R>& v, R t) {
    const R ur = u.re;
    u.re = qk_fma(-t, v.re, ur);
    v.re = qk_fma(t, ur, v.re);
}
template <typename R> QK_HD Cplx<R> lin2(const Cplx<R>& u, const Cplx<R>& v, R a, R b) {
    Cplx<R> o;
    o.re = qk_fma(a, u.re, b * v.re);
    o.im = qk_fma(a, u.im, b * v.im);
    return o;
}
template <typename R> QK_HD Real<R> lin2(const Real<R>& u, const Real<R>& v, R a, R b) {
    Real<R> o;
    o.re = qk_fma(a, u.re, b * v.re);
    return o;
}
template <typename R> QK_HD void fma_amp(Cplx<R>& acc, R c, const Cplx<R>& z) {
    acc.re = qk_fma(c, z.re, acc.re);
    acc.im = qk_fma(c, z.im, acc.im);
}
template <typename R> QK_HD void fma_amp(Real<R>& acc, R c, const Real<R>& z) { acc.re = qk_fma(c, z.re, acc.re); }

// Same contract as evolve_blocks, accumulating into acc:
//   CHEB applications 1 .. DT-1: full scaled passes on both f_w halves              8 instr / 16 flops
//   CHEB application DT (pruned to f_x = 0, with gamma^D J^D):  alpha u + beta v    8 instr / 12 flops
//   MUL (pruned to (0,0)) fused with the read-out sum: acc += cw lo0 - sw lo2       4 instr /  8 flops
// (complex amplitudes; half of that for the real-only representation)
template <class A, typename R, int U, int DT>
QK_HD void evolve_blocks_tan(const A (&init)[4], const R (&t)[U], const R (&al)[U], const R (&be)[U], const R (&cw)[U],
                             const R (&sw)[U], A& acc) {
    A v[U][4];
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        QK_UNROLL
        for (int q = 0; q < 4; ++q) v[u][q] = init[q];
    }
    QK_UNROLL
    for (int r = 0; r + 1 < DT; ++r) {
        QK_UNROLL
        for (int u = 0; u < U; ++u) {
            rot_tan(v[u][0], v[u][1], t[u]);
            rot_tan(v[u][2], v[u][3], t[u]);
        }
    }
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        const A lo0 = lin2(v[u][0], v[u][1], al[u], be[u]);
        const A lo2 = lin2(v[u][2], v[u][3], al[u], be[u]);
        fma_amp(acc, cw[u], lo0);
        fma_amp(acc, -sw[u], lo2);
    }
}

#if defined(__CUDACC__)
// result store: to every listed buffer (local + NVLink peers), or once through the multicast mapping
__device__ __forceinline__ void store_result(const BlockParams& p, long long idx, double val) {
    if (p.mc_out != nullptr) {
        asm volatile("multimem.st.weak.global.f64 [%0], %1;" ::"l"(p.mc_out + idx), "d"(val) : "memory");
    } else {
        p.outs[0][idx] = val;
#pragma unroll 1
        for (int q = 1; q < p.n_out; ++q) p.outs[q][idx] = val;     // NVLink peers (fused gather)
    }
}

// keep a kernel parameter in a register: without this the compiler re-reads the prepared block state from the
// constant bank inside the pass loop (LDC, a long-scoreboard load: it was the top stall of the window kernel)
__device__ __forceinline__ void keep_in_register(double& v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void keep_in_register(float& v) { asm volatile("" : "+f"(v)); }

template <class A> __device__ __forceinline__ A shfl_xor_amp(const A& a, int m) {
    A r;
    r.re = __shfl_xor_sync(0xffffffffu, a.re, m);
    if constexpr (A::is_complex) r.im = __shfl_xor_sync(0xffffffffu, a.im, m);
    return r;
}

// SIMPLE: every lane owns one whole output row of its sample (G_r = 1 and a single row step - the
// layout of every small BASELINE layer): no row loop and no shuffle step in the per-sample code.
template <class A, typename R, int U, int SU, int MODE, int NT, int MINB, bool SIMPLE, int DT>
__global__ void __launch_bounds__(NT, MINB) qkan_block_kernel(const BlockParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = p.G, G_r = p.G_r, G_k = p.G_k;
    const int SPC = p.SPC;                                   // samples in flight per CTA
    const int tile = p.tile;                                 // samples per x tile
    const int RB = p.row_bytes;                              // cs row stride: N entries + the dummy (+ padding)
    constexpr size_t ENTB = sizeof(CS<R>);                   // bytes per entry
    // smem: xs[2] (TMA destinations: raw x rows, two tiles in flight) | cs (rotation entries of the current tile, plus
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
    // this CTA's slice of the batch: samples [base, bend), walked tile by tile (s_tot < 0: the A/B alternative -
    // tiles of the whole batch dealt round-robin to the CTAs)
    const bool strided = p.s_tot < 0;
    const long long base = strided ? 0 : ((long long)blockIdx.x * p.s_tot / gridDim.x) * SPC;
    const long long bend_raw = strided ? p.B : (((long long)blockIdx.x + 1) * p.s_tot / gridDim.x) * SPC;
    const long long bend = bend_raw < p.B ? bend_raw : p.B;
    const long long it_step = strided ? (long long)gridDim.x : 1;
    const long long n_it = (bend - base + tile - 1) / tile;
    const CS<R>* __restrict__ cstab = reinterpret_cast<const CS<R>*>(p.cstab);
    const int* __restrict__ xotab = p.xotab;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    // the dummy entries never change
    for (int i = tid; i < tile; i += NT) {
        CS<R> e;
        e.c = R(0); e.s = R(1);
        *reinterpret_cast<CS<R>*>(cs + (size_t)i * RB + p.N * ENTB) = e;
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
#ifdef QKAN_PIN_INIT       // A/B-tested on this kernel (its tables are L1 resident): no measurable difference, left off
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        keep_in_register(init[q].re);
        if constexpr (A::is_complex) keep_in_register(init[q].im);
    }
#endif
    // per-lane view of the slot tables: entry of (bi, pi, u) sits ((bi * passes + pi) * U + u) * G after `g`
    R cw[U], sw[U];
    int xoff[U], deg[U];
    auto unpack = [&](int u, const CS<R>& q, int xo) {
        cw[u] = q.c; sw[u] = q.s;
        if constexpr (MODE == 1) { xoff[u] = xo & 0xFFFFFF; deg[u] = xo >> 24; }
        else { xoff[u] = xo; deg[u] = 0; }
    };
    // entries of the first pass: resident for the whole launch, so a new sample never waits for a table load
    CS<R> q0[U];
    int x0[U];
    QK_UNROLL
    for (int u = 0; u < U; ++u) { q0[u] = cstab[(size_t)u * G + g]; x0[u] = xotab[(size_t)u * G + g]; }

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
        // ChebyshevStep.py:46-49), clip (:52) and the rotation pair cos(theta/2) = x,
        // sin(theta/2) = sqrt(1 - x^2) - no arccos is ever needed
        unsigned bad = 0;
        {
            const int n_in = nsamp * p.N;
            int row = pre_row0, n = pre_n0;
            for (int e = tid; e < n_in; e += NT) {
                const double v = xs[e];
                if (!(fabs(v) <= 1.0 + 1e-8)) ++bad;
                const R c = clip_unit<R>(v);
                CS<R> en;
                en.c = c;
                en.s = qk_sqrt((R(1) - c) * (R(1) + c));
                *reinterpret_cast<CS<R>*>(cs + (size_t)row * RB + n * ENTB) = en;
                n += pre_dn;
                row += pre_dr;
                if (n >= p.N) { n -= p.N; ++row; }
            }
        }
        if (bad) atomicAdd(p.oor, (unsigned long long)bad);
        __syncthreads();                                      // cs complete, xs[buf] free again
        const long long nxt = it + 2 * it_step;
        if (nxt < n_it) issue_x(nxt, buf);                    // overlaps with the compute of this and the next tile

        // SU samples per lane at a time (the lane's slots of SU consecutive sub-iterations): they share
        // every table entry, so the loads, pointer bumps and per-sample set-up are paid once per SU samples
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
                row[j] = csrow + j * row_stride;      // idle slots of a ragged tile evolve a stale row of the tile; nothing is stored
            }
            // stream the lane's slots with one running pointer; the next pass's entries are fetched while
            // the current pass is evolved (the tables end with one pass of padding slots)
            const CS<R>* cp = cstab + g;
            const int* xp = xotab + g;
            CS<R> qn[U];
            int xn[U];
            QK_UNROLL
            for (int u = 0; u < U; ++u) { qn[u] = q0[u]; xn[u] = x0[u]; }
            auto run_row = [&](A (&acc)[SU]) {
                QK_UNROLL
                for (int j = 0; j < SU; ++j) set_amp(acc[j], 0.0);
                for (int pi = 0; pi < p.passes; ++pi) {
                    QK_UNROLL
                    for (int u = 0; u < U; ++u) unpack(u, qn[u], xn[u]);
                    cp += (size_t)U * G;
                    xp += (size_t)U * G;
                    QK_UNROLL
                    for (int u = 0; u < U; ++u) { qn[u] = cp[(size_t)u * G]; xn[u] = xp[(size_t)u * G]; }
                    {
                        R cx[SU][U], sx[SU][U];
                        QK_UNROLL
                        for (int j = 0; j < SU; ++j) {
                            QK_UNROLL
                            for (int u = 0; u < U; ++u) {
                                const CS<R> e = *reinterpret_cast<const CS<R>*>(row[j] + xoff[u]);
                                cx[j][u] = e.c; sx[j][u] = e.s;
                            }
                        }
                        QK_UNROLL
                        for (int j = 0; j < SU; ++j) {
                            const A part = evolve_blocks<A, R, U, MODE, DT>(init, cx[j], sx[j], cw, sw, deg, p.D);
                            add_amp(acc[j], part);
                        }
                    }
                }
            };
            auto write_row = [&](const A& acc, int b, int j) {
                const double val = (double)acc.re * p.out_scale;
                store_result(p, o + j * out_stride + b, val);
                if (p.amps) {
                    Cplx<R> z;
                    z.re = (R)((double)acc.re * p.amp_scale);
                    if constexpr (A::is_complex) z.im = (R)((double)acc.im * p.amp_scale);
                    else z.im = R(0);
                    reinterpret_cast<Cplx<R>*>(p.amps)[oa + j * out_stride + b] = z;
                }
            };
            A acc[SU];
            if constexpr (SIMPLE) {
                // every lane owns one whole output row (G_r = 1, one row step): UNPREPARE + SUM +
                // post-selection is the lane's running sum, no cross-lane step
                run_row(acc);
                QK_UNROLL
                for (int j = 0; j < SU; ++j)
                    if (valid[j] && k < p.K) write_row(acc[j], k, j);
            } else {
                for (int b = k; b < p.brows * G_k; b += G_k) {
                    run_row(acc);
                    // UNPREPARE (H on deg) + SUM (H on a) + post-selection deg = a = 0: the sum over the
                    // row's blocks, finished across the G_r lanes with an xor butterfly
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) {
                        for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc[j], shfl_xor_amp(acc[j], m));
                        if (valid[j] && r == 0 && b < p.K) write_row(acc[j], b, j);
                    }
                }
            }
        }
        __syncthreads();                                      // everyone done with cs before the next pre-pass
    }
}


template <typename R>
__global__ void qkan_prepare_block_tables_kernel(const double* W, int N, int K, int D, int U, int passes, int g_r_log2,
                                                 int g_k_log2, int paper, int x_entry_bytes, int window, long long slots_total,
                                                 CS<R>* cstab, int* xotab, unsigned long long* bad_weights) {
    const long long slot = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (slot >= slots_total) return;
    fill_block_slot<R>(slot, W, N, K, D, U, passes, g_r_log2, g_k_log2, paper, cstab, xotab, x_entry_bytes, window);
    // |w| <= 1 is required for the rotation to exist (MulStep.py:36-37): check each weight once
    if (slot < (long long)N * K * (D + 1)) {
        const double w = W[slot];
        if (!(fabs(w) <= 1.0)) atomicAdd(bad_weights, 1ull);
    }
}

struct BlockKernelInfo {
    int amp, mode, U, NT, MINB;
    int SU;                     // samples per lane at a time
    int DT;                     // 0 = any D (run-time loop), else only for D == DT
    int tan;                    // scaled-rotation form: cs rows hold (t, alpha, beta) triples (qkan_amajor.cuh)
    int window;                 // window kernel (wide input rows): entries are built per row step
    int amajor;                 // a-major tables (qkan_amajor.cuh): D + 1 SELECT entries per (row step, pass, lane)
    int amp_bytes;              // a-major kernels: sizeof(amplitude), the entry size of the cs planes
    int direct;                 // direct kernel (qkan_amajor.cuh): every output row reads one input element, no shared memory
    int elem;                   // element-owner kernel (qkan_amajor.cuh): wide input rows, lanes own input elements, no shared memory
    int is_default;
    cudaError_t (*launch)(const BlockParams&, int g, int sm_count, cudaStream_t, int* grid_out, int* smem_out);
};

template <class A, typename R, int U, int SU, int MODE, int NT, int MINB, bool SIMPLE, int DT>
cudaError_t launch_block_impl(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    auto kern = qkan_block_kernel<A, R, U, SU, MODE, NT, MINB, SIMPLE, DT>;
    BlockParams p = p0;
    const int SPC = NT / G;
    p.row_bytes = cs_row_stride(p.N, G, (int)sizeof(CS<R>)) * (int)sizeof(CS<R>);
    // wide rows: staging the raw x twice more than doubles the shared memory per sample and would
    // halve the resident warps; the per-tile compute is long, so the pre-pass reads global memory directly
    p.direct_x = ((size_t)SPC * p.N * 16 > 16 * 1024) ? 1 : 0;
    auto smem_for = [&](int sub) {
        const size_t tile = (size_t)SPC * sub;
        const size_t xs = p.direct_x ? 0 : 2 * ((tile * p.N + 1) & ~(size_t)1) * sizeof(double);
        const size_t cs = ((tile + (size_t)(SU - 1) * SPC) * (size_t)p.row_bytes + 15) & ~(size_t)15;
        return xs + cs + 16;
    };
    int sub = (int)(8192 / ((size_t)SPC * p.N * 8));          // about 8 KiB of x per tile ...
    const int sub_cs = (int)(24576 / ((size_t)SPC * p.row_bytes));   // ... and at most 24 KiB of rotation pairs
    if (sub > sub_cs) sub = sub_cs;
    if (const char* e = getenv("QKAN_BLOCK_SUB")) sub = atoi(e);   // tuning aid
    if (sub > 32) sub = 32;
    // a lane takes SU sub-iterations at a time: a tile of an odd number of them would leave a sample slot idle
    auto round_su = [](int v) { v -= v % SU; return v < SU ? SU : v; };
    sub = round_su(sub);
    if (smem_for(sub) > 200 * 1024) return cudaErrorInvalidConfiguration;
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
    if (const char* e = getenv("QKAN_BLOCK_STRIDED")) strided = atoi(e) != 0;   // A/B aid
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
template <class A, typename R, int U, int SU, int MODE, int NT, int MINB, int DT>
cudaError_t launch_block(const BlockParams& p, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (DT > 0 && p.D != DT) return cudaErrorInvalidValue;
    if (p.g_r_log2 == 0 && p.brows == 1)
        return launch_block_impl<A, R, U, SU, MODE, NT, MINB, true, DT>(p, G, sm_count, stream, grid_out, smem_out);
    return launch_block_impl<A, R, U, SU, MODE, NT, MINB, false, DT>(p, G, sm_count, stream, grid_out, smem_out);
}

template <class A> struct AmpId;

template <class A, typename R, int U, int SU, int MODE, int NT, int MINB, int DT>
BlockKernelInfo make_block_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = MODE; k.U = U; k.SU = SU; k.NT = NT; k.MINB = MINB; k.DT = DT; k.is_default = is_default;
    k.tan = 0; k.window = 0; k.amajor = 0; k.direct = 0; k.elem = 0; k.amp_bytes = (int)sizeof(A);
    k.launch = &launch_block<A, R, U, SU, MODE, NT, MINB, DT>;
    return k;
}
#endif  // __CUDACC__

}  // namespace qkan
