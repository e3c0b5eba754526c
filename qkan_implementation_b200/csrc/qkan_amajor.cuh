// qkan_amajor.cuh - the default forward kernels of the block engine (compat mode, 1 <= D <= 16).
//
// Same circuit as qkan_block.cuh, simulated block by block in the scaled-rotation form, but
// using two more facts about the circuit's STRUCTURE (nothing that depends on the input values):
//
//  (1) CHEB acts on f_x only, and PREPARE leaves every index register in a product state.  Until SELECT the
//      statevector is therefore  sum_(a,b) |a,b> (x) (block state of (a,b)) (x) |+>_deg : the D + 1 degree copies
//      of an (a, b) block hold the SAME four amplitudes, so they are evolved through the CHEB sequence once, not
//      D + 1 times.  (With the rotation entry in registers the D + 1 evolutions are literally common
//      subexpressions - ptxas merged them in a kernel that spelled out every block - so the code says it.)
//  (2) The CHEB multiplexor UCRy(theta_x) is controlled by (a, b) but its angle table has only N distinct entries:
//      theta_x[a, b] = theta(x[(a + N b) / K])  (np.repeat dilation, ChebyshevStep.py:64, against the column-major
//      SUM reshape, QKANLayer.py:132).  Blocks that share the entry share the evolution, so the CHEB sequence is
//      run once per INPUT ELEMENT: the pre-pass over a tile's inputs produces, per element, the two f_x = 0 block
//      amplitudes after CHEB, (lo0, lo2) = (f_w = 0, f_w = 1): 8 (D - 1) FMA + 4 MUL + 4 FMA per element.
//
// SELECT (MUL) is controlled by (a, b, deg): every one of the N K (D + 1) blocks gets its own rotation
// (cos, sin)(theta_w / 2), applied to its copy of (lo0, lo2) and fused with the read-out sum: 4 FMA per block.
// No weight is pre-summed and nothing state dependent is skipped (the f_w = 1 half and the imaginary parts are
// zero at run time and are evolved like any other amplitude; the prepared block state is a kernel parameter).
// Per sample: 8 D N + 4 N K (D + 1) FP instructions, against N K (D + 1)(8 D + 4) when every block is evolved on
// its own (round 1; qkan_kernel_info.flops_per_block_basis keeps that count for comparison).
//
// Mapping: one sample = G = G_k * G_r lanes of one warp.  Lane (k, r) owns output rows b = bi G_k + k and, in
// row b, the summed indices a = pi G_r + r, pi = 0 .. passes - 1 (passes = ceil(N / G_r)); it applies SELECT to
// all D + 1 degree copies of (a, b).  UNPREPARE + SUM + post-selection on deg = a = 0 is the lane's running sum,
// finished across the G_r lanes of the row by an xor butterfly.  A lane re-reads (lo0, lo2) from shared memory only
// when the input element changes from one a to the next (never, for K a multiple of N).
//
// Tables (built once per set_weights by qkan_prepare_amajor_tables_kernel):
//     step = ((bi * passes + pi) << g_log2) + g,   g = (k << g_r_log2) | r
//     wtab[(((bi * passes + pi) * (D + 1) + d) << g_log2) + g]
//                              = (cos, sin)(theta_w / 2) of block (a, b, d): W[d][a + N b] (MulStep.py:69); the lane is
//                                the fastest index, so one warp load reads one contiguous run of G entries (one 128-byte
//                                line for G <= 8 in FP64: the L1 pipe takes about two cycles per line a load touches, and
//                                a lane-major table - D + 1 entries of a lane contiguous - made every load touch G lines,
//                                which bound the deep sequences, profiles/r02c_tune_*.jsonl)
//     xotab[step]              = byte offset of lo0 of x[(a + N b) / K] in the sample's cs row (window kernel:
//                                relative to the row step's input window); lo2 sits `plane` bytes further
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

// table entries of one (bi, pi, lane) step: D + 1 SELECT rotations and the offset of the element's lo0
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
        wtab[((((step >> g_log2) * (D + 1)) + d) << g_log2) + g] = q;
    }
}

// steps of the tables: every (row step, pass, lane) plus one pass of padding (the kernels prefetch one pass ahead)
inline long long amajor_steps(const BlockLayout& lay) {
    return ((long long)lay.brows * lay.passes + 1) << (lay.g_r_log2 + lay.g_k_log2);
}

// A sample's cs row: two planes of n1 = N + 1 amplitudes (lo0[0 .. N], lo2[0 .. N]; index N = the dummy, x = 0),
// `amp_bytes` each.  Returns the row stride in amplitudes (>= 2 n1; every amplitude stays naturally aligned).  One shared-memory
// phase serves 128 bytes = P lanes; with G < P lanes per sample a phase spans P / G consecutive sample rows whose
// lanes read G consecutive amplitudes of a plane, so the stride is chosen to spread them over the P slots.
inline int amajor_row_amps(int n1, int G, int amp_bytes) {
    const int P = 128 / amp_bytes;
    const int need = 2 * n1;
    if (G >= P || n1 > 2 * P) return need;
    int best = need, best_worst = 1 << 30;
    for (int rs = need; rs < need + P; ++rs) {
        int cnt[32] = {0};
        int worst = 0;
        for (int lane = 0; lane < P; ++lane) {
            const int w = ((lane / G) * rs + (lane % G) % n1) % P;
            if (++cnt[w] > worst) worst = cnt[w];
        }
        if (worst < best_worst) { best_worst = worst; best = rs; }
    }
    return best;
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

// CHEB on one input element: the block state `init` through D applications of Ry(theta_x), cos(theta_x / 2) = c;
// returns the f_x = 0 amplitudes of the two f_w halves.  Two forms, chosen per compile-time degree:
//
//   sin-weighted basis (D <= SW_FORM_MAX_DT).  The f_x = 1 amplitude of every pair is stored multiplied by
//       s = sin(theta_x / 2):  (u, w) = (amp[f_x = 0], s * amp[f_x = 1]).  In that basis Ry(theta_x) reads
//           u' = c u - w,        w' = s^2 u + c w,        s^2 = 1 - c^2    (one FMA, exact to half an ulp)
//       so the rotation entry is (c, s^2): no square root, no reciprocal, no case selection.  A full pass over a
//       block is 4 MUL + 8 FMA (complex amplitudes, both f_w halves); the last pass, pruned to its f_x = 0 outputs,
//       is 4 FMA - the post-selected amplitude is a u component, so the basis scale never has to be undone.  The
//       prepared block state has no f_x = 1 component (PREPARE leaves f_x in |0>), so it is the same vector in this
//       basis; BlockParams::init is read as stored amplitudes.  (A diagonal change of basis of the simulated state -
//       the same device as deferring gamma^D in the scaled form or the 1/sqrt(2) of every Hadamard to the read-out.)
//       Per element: 1 + 12 (D - 1) + 4 FP64 instructions.
//   scaled form (D > SW_FORM_MAX_DT): D - 1 full passes M(t) (8 FMA), then the pruned pass alpha u + beta v with
//       gamma^D and the quarter turns (4 MUL + 4 FMA); the conversion of c into (t, alpha, beta) costs one reciprocal
//       square root and about 24 FP64 instructions + 14 selects per element: 24 + 8 (D - 1) + 8.
// The plain (cos, sin) form of round 2's first session (D <= 2: one square root + 16 (D - 1) + 8) is superseded by the
// sin-weighted basis at every degree (profiles/r02p_bench_cs_form_d3.json has its measurements).
#ifndef QKAN_SW_FORM_MAX_DT
#define QKAN_SW_FORM_MAX_DT 8
#endif
constexpr int SW_FORM_MAX_DT = QKAN_SW_FORM_MAX_DT;
constexpr bool cheb_uses_sw_form(int DT) { return DT <= SW_FORM_MAX_DT; }

// one full pass in the sin-weighted basis: (u, w) <- (c u - w, s2 u + c w)
template <typename R> QK_HD void rot_sw(Cplx<R>& u, Cplx<R>& w, R c, R s2) {
    const R ur = u.re, ui = u.im;
    u.re = qk_fma(c, ur, -w.re);
    u.im = qk_fma(c, ui, -w.im);
    w.re = qk_fma(s2, ur, c * w.re);
    w.im = qk_fma(s2, ui, c * w.im);
}
template <typename R> QK_HD void rot_sw(Real<R>& u, Real<R>& w, R c, R s2) {
    const R ur = u.re;
    u.re = qk_fma(c, ur, -w.re);
    w.re = qk_fma(s2, ur, c * w.re);
}
// its f_x = 0 output only: c u - w
template <typename R> QK_HD Cplx<R> rot_sw_lo(const Cplx<R>& u, const Cplx<R>& w, R c) {
    Cplx<R> o;
    o.re = qk_fma(c, u.re, -w.re);
    o.im = qk_fma(c, u.im, -w.im);
    return o;
}
template <typename R> QK_HD Real<R> rot_sw_lo(const Real<R>& u, const Real<R>& w, R c) {
    Real<R> o;
    o.re = qk_fma(c, u.re, -w.re);
    return o;
}

template <class A, typename R, int DT>
QK_HD void cheb_element(const A (&init)[4], R c, A& lo0, A& lo2) {
    A v[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) v[q] = init[q];
    if constexpr (cheb_uses_sw_form(DT)) {
        const R s2 = qk_fma(-c, c, R(1));
        QK_UNROLL
        for (int r = 0; r + 1 < DT; ++r) {
            rot_sw(v[0], v[1], c, s2);
            rot_sw(v[2], v[3], c, s2);
        }
        lo0 = rot_sw_lo(v[0], v[1], c);
        lo2 = rot_sw_lo(v[2], v[3], c);
    } else {
        const TanEntry<R> e = tan_entry<R>(c, DT);
        QK_UNROLL
        for (int r = 0; r + 1 < DT; ++r) {
            rot_tan(v[0], v[1], e.t);
            rot_tan(v[2], v[3], e.t);
        }
        lo0 = lin2(v[0], v[1], e.al, e.be);
        lo2 = lin2(v[2], v[3], e.al, e.be);
    }
}

// SELECT on the D + 1 degree copies of SU samples' (a, b) block, fused with the read-out sum:
// acc += cos(theta_w / 2) lo0 - sin(theta_w / 2) lo2   (the (f_x, f_w) = (0, 0) output of Ry(theta_w) on f_w)
// wp -> the lane's entry of degree 0; consecutive degrees are G entries apart
template <class A, typename R, int SU, int DT>
QK_HD void select_blocks(const A (&lo0)[SU], const A (&lo2)[SU], const CS<R>* __restrict__ wp, int G, A (&acc)[SU]) {
    QK_UNROLL
    for (int d = 0; d <= DT; ++d) {
        const CS<R> q = wp[(size_t)d * G];
        QK_UNROLL
        for (int j = 0; j < SU; ++j) {
            fma_amp(acc[j], q.c, lo0[j]);
            fma_amp(acc[j], -q.s, lo2[j]);
        }
    }
}

// direct kernel applies: one row step, one lane per row, each row reads a single input element
inline bool amajor_direct_ok(int N, int K, const BlockLayout& lay) {
    return K % N == 0 && lay.g_r_log2 == 0 && lay.brows == 1 && lay.efficiency == 1.0;
}

// ---------------------------------------------------------------------------------------------------------------
// Element-owner walk (wide input rows, e.g. N784 K10): output row b reads the inputs n_first(b) .. n_last(b),
// n = (a + N b) / K, about N / K of them, and input n feeds the K consecutive summed indices a = n K - N b + j,
// j < K.  Lane (k, r) of a sample owns the elements e = pi G_r + r of its row b = bi G_k + k: it loads x[n] itself,
// runs the CHEB sequence in registers and applies SELECT to the element's K (D + 1) blocks.  Nothing is staged in
// shared memory, so SU samples per lane cost registers only, and every SELECT entry a warp loads serves
// (32 / G) SU samples - what keeps the L1 pipe (one 128-byte wavefront per cycle) behind the FP64 pipe.
// Tables:   xe[(bi * passes + pi) * G + g]                       = n | (count << 30) | (live << 31 as sign: -1 = padding)
//           we[(((bi * passes + pi) * K + j) * (D + 1) + d) * G + g] = (cos, sin)(theta_w / 2) of block (a, b, d)
struct ElemLayout {
    int g_r_log2, g_k_log2, passes, brows;
    double efficiency;
};
QK_HD void elem_row_range(int N, int K, long long b, long long* n_first, long long* n_last) {
    *n_first = (b * N) / K;
    *n_last = (b * N + N - 1) / K;
}
inline ElemLayout plan_elem_layout(int N, int K, int min_g_log2 = 0) {
    ElemLayout best{};
    double best_score = -1.0;
    long long wmax = 1;
    for (long long b = 0; b < K; ++b) {
        long long f, l;
        elem_row_range(N, K, b, &f, &l);
        if (l - f + 1 > wmax) wmax = l - f + 1;
    }
    // rows one after the other (G_k = 1: the most samples per warp share a table load); the lanes of a row split its
    // elements: the widest split up to 16 lanes (16 consecutive inputs = one 128-byte line per load) that keeps the
    // padding of the last pass below 10 % of the best split
    double eff_max = 0.0;
    for (int gr = 0; gr <= 5; ++gr) {
        const long long G_r = 1ll << gr, passes = (wmax + G_r - 1) / G_r;
        const double eff = (double)N / (double)(passes * G_r * K);      // live blocks / issued block slots of a row
        if (eff > eff_max) eff_max = eff;
    }
    int force_gr = -1;
    if (const char* e = getenv("QKAN_ELEM_GR")) force_gr = atoi(e);      // tuning aid: log2 of the lanes per row
    for (int gr = 0; gr <= 5; ++gr) {
        if (gr < min_g_log2 && gr < 5) continue;
        if (force_gr >= 0 && gr != force_gr) continue;
        const long long G_r = 1ll << gr, passes = (wmax + G_r - 1) / G_r;
        const double eff = (double)N / (double)(passes * G_r * K);
        if (force_gr < 0 && best_score >= 0.0 && (gr > 4 || eff < 0.9 * eff_max)) continue;
        best_score = eff;
        best.g_r_log2 = gr; best.g_k_log2 = 0; best.passes = (int)passes; best.brows = K; best.efficiency = eff;
    }
    return best;
}
inline long long elem_steps(const ElemLayout& lay) { return ((long long)lay.brows * lay.passes + 1) << (lay.g_r_log2 + lay.g_k_log2); }

template <typename R>
QK_HD void fill_elem_step(long long step, const double* W, int N, int K, int D, int passes, int brows, int g_r_log2, int g_k_log2,
                          CS<R>* we, int* xe) {
    const int g_log2 = g_r_log2 + g_k_log2;
    const int g = (int)(step & ((1ll << g_log2) - 1));
    const long long t = step >> g_log2;
    const int pi = (int)(t % passes);
    const long long bi = t / passes;
    const int k = g >> g_r_log2, r = g & ((1 << g_r_log2) - 1);
    const long long b = (bi << g_k_log2) + k;
    long long nf = 0, nl = -1, nl_prev = -1;
    if (bi < brows && b < K) {
        elem_row_range(N, K, b, &nf, &nl);
        if (b > 0) { long long f2; elem_row_range(N, K, b - 1, &f2, &nl_prev); }
    }
    const long long n = nf + ((long long)pi << g_r_log2) + r;
    const bool live = bi < brows && b < K && n <= nl;
    // a padding step evaluates element 0 (any valid address) and multiplies it by theta = pi rotations: adds exactly 0
    xe[step] = live ? (int)n | (n > nl_prev ? (1 << 30) : 0) : -1;      // bit 30: the first row that reads the element counts its range violation
    if (bi >= brows) return;                                            // the padding pass has no SELECT entries
    for (int j = 0; j < K; ++j) {
        const long long a = n * K - (long long)N * b + j;
        const bool on = live && a >= 0 && a < N;
        for (int d = 0; d <= D; ++d) {
            CS<R> q;
            q.c = R(0); q.s = R(1);
            if (on) {
                const R w = (R)W[(long long)d * N * K + a + (long long)N * b];
                q.c = w;
                q.s = qk_sqrt((R(1) - w) * (R(1) + w));
            }
            we[((((t * K) + j) * (D + 1) + d) << g_log2) + g] = q;
        }
    }
}

#if defined(__CUDACC__)
// |v| >= 1, infinite or NaN, decided on the high word alone (one LOP3 + one ISETP on the integer pipe; the FP64 pipe is the
// scarce one).  Inputs strictly inside (-1, 1) - all of a normalised batch - need neither the clip nor the range count
// (ChebyshevStep.py:46-52); the rare others take the exact path.
__device__ __forceinline__ bool at_or_beyond_unit(double v) { return (__double2hiint(v) & 0x7fffffff) >= 0x3ff00000; }

// result store.  plain: one local buffer, no amplitudes (a single uniform branch in the hot path)
template <class A, typename R>
__device__ __forceinline__ void amajor_store(const BlockParams& p, const A& acc, double* __restrict__ out_t, void* amps_t, int o) {
    const double val = (double)acc.re * p.out_scale;
    if (p.plain) {
        out_t[o] = val;
        return;
    }
    store_result(p, (long long)(out_t - p.outs[0]) + o, val);
    if (amps_t) {
        Cplx<R> z;
        z.re = (R)((double)acc.re * p.amp_scale);
        if constexpr (A::is_complex) z.im = (R)((double)acc.im * p.amp_scale);
        else z.im = R(0);
        reinterpret_cast<Cplx<R>*>(amps_t)[o] = z;
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
    const int RB = p.row_bytes;                              // cs row stride
    const int plane = p.plane_bytes;                         // lo2 plane offset inside a row
    constexpr int D1 = DT + 1;
    // smem: xs[2] (TMA destinations: raw x rows, two tiles in flight) | cs ((lo0, lo2) of the current tile's inputs, plus
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
    const int it_step = strided ? (int)gridDim.x : 1;
    const int n_it = (int)((bend - base + tile - 1) / tile);
    const CS<R>* __restrict__ wtab = reinterpret_cast<const CS<R>*>(p.cstab);
    const int* __restrict__ xotab = p.xotab;

    A init[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        init[q].re = (R)p.init[2 * q];
        if constexpr (A::is_complex) init[q].im = (R)p.init[2 * q + 1];
    }
    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    {   // the dummy entries (x = 0) never change
        A d0, d2;
        cheb_element<A, R, DT>(init, R(0), d0, d2);
        for (int i = tid; i < tile + (SU - 1) * SPC; i += NT) {
            *reinterpret_cast<A*>(cs + (size_t)i * RB + p.N * sizeof(A)) = d0;
            *reinterpret_cast<A*>(cs + (size_t)i * RB + plane + p.N * sizeof(A)) = d2;
        }
    }
    __syncthreads();

    auto tile_samples = [&](int i) -> int {
        const long long left = bend - (base + (long long)i * tile);
        return left < tile ? (int)left : tile;
    };
    // stage the x rows of a tile into buffer b: one 1-D TMA bulk copy when the tile is 16-byte
    // granular, plain coalesced loads otherwise (ragged tail, odd N)
    auto issue_x = [&](int i, int b) {
        if (p.direct_x) return;
        const unsigned bytes = (unsigned)tile_samples(i) * (unsigned)p.N * 8u;
        const double* src = p.x + (base + (long long)i * tile) * p.N;
        double* dst = xs0 + (size_t)b * xs_doubles;
        if (p.tma_ok && (bytes & 15u) == 0) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&mbar[b], bytes);
                tma_load_1d(dst, src, bytes, &mbar[b]);
            }
        } else {
            for (int q = tid; q < (int)(bytes >> 3); q += NT) dst[q] = src[q];
        }
    };

    const int x0 = xotab[g];                                  // first step's offset: resident for the whole launch
    const size_t wstep = (size_t)G * D1;                      // table entries between consecutive passes of a lane

    // two tiles in flight: tile i is consumed while tiles i+1 and (after its pre-pass) i+2 are loading
    int it = strided ? (int)blockIdx.x : 0;
    unsigned phase0 = 0, phase1 = 0;
    int buf = 0;
    if (it < n_it) issue_x(it, 0);
    if (it + it_step < n_it) issue_x(it + it_step, 1);
    __syncthreads();
    // pre-pass walk: thread tid takes inputs tid, tid + NT, ... of the tile; its (row, n) advances by
    // (NT / N, NT % N) per step, so the loop has no division
    const int pre_row0 = tid / p.N, pre_n0 = tid - pre_row0 * p.N;
    const int pre_dr = NT / p.N, pre_dn = NT - pre_dr * p.N;
    char* const pre_dst0 = cs + (size_t)pre_row0 * RB + pre_n0 * (int)sizeof(A);
    const int pre_step = pre_dr * RB + pre_dn * (int)sizeof(A);       // bytes per walk step without the row wrap
    const int pre_wrap = RB - p.N * (int)sizeof(A);                   // extra bytes when n wraps into the next row

    const int row_stride = SPC * RB;                          // bytes between consecutive sub-iterations
    const int out_stride = SPC * p.K;
    const char* const csrow0 = cs + (size_t)slot * RB;

    for (; it < n_it; it += it_step, buf ^= 1) {
        const int nsamp = tile_samples(it);
        if (!p.direct_x && p.tma_ok && ((nsamp * p.N) & 1) == 0) {
            if (buf == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
        }
        const long long s0 = base + (long long)it * tile;
        const double* xs = p.direct_x ? p.x + s0 * p.N : xs0 + (size_t)buf * xs_doubles;

        // pre-pass over the raw inputs of the tile: range count (the reference prints a warning,
        // ChebyshevStep.py:46-49), clip (:52) and the CHEB sequence of the element: cos(theta/2) = x, no arccos
        unsigned bad = 0;
        {
            const int n_in = nsamp * p.N;
            int n = pre_n0;
            char* dst = pre_dst0;
            for (int e = tid; e < n_in; e += NT) {
                double v = xs[e];
                if (at_or_beyond_unit(v)) {                   // |v| >= 1 or NaN: the exact test and the clip
                    if (!(fabs(v) <= 1.0 + 1e-8)) ++bad;
                    v = (double)clip_unit<double>(v);
                }
                A lo0, lo2;
                cheb_element<A, R, DT>(init, (R)v, lo0, lo2);
                *reinterpret_cast<A*>(dst) = lo0;
                *reinterpret_cast<A*>(dst + plane) = lo2;
                n += pre_dn;
                dst += pre_step;
                if (n >= p.N) { n -= p.N; dst += pre_wrap; }
            }
        }
        if (bad) atomicAdd(p.oor, (unsigned long long)bad);
        __syncthreads();                                      // cs complete, xs[buf] free again
        if (it + 2 * it_step < n_it) issue_x(it + 2 * it_step, buf);   // overlaps with the compute of this and the next tile

        // SU samples per lane at a time (the lane's slots of SU consecutive sub-iterations): they share every
        // SELECT entry and the per-pass bookkeeping
        double* const out_t = p.outs[0] + (p.row0 + s0) * p.K;
        void* const amps_t = p.amps ? (void*)(reinterpret_cast<Cplx<R>*>(p.amps) + s0 * p.K) : nullptr;
        const char* csrow = csrow0;
        int o = slot * p.K;
        // whole sub-iterations (every lane of a sample group runs: the butterfly needs all of them)
        const int ls_end = (nsamp + SPC - 1) & ~(SPC - 1);
        for (int ls = slot; ls < ls_end; ls += SU * SPC, csrow += SU * row_stride, o += SU * out_stride) {
            const char* row[SU];
            QK_UNROLL
            for (int j = 0; j < SU; ++j) row[j] = csrow + j * row_stride;      // idle slots of a ragged tile evolve a stale / slack row; nothing is stored
            const CS<R>* wp = wtab + g;
            const int* xp = xotab + g;
            int xo = x0, xprev = -1;
            A acc[SU], lo0[SU], lo2[SU];
            auto run_row = [&]() {
                QK_UNROLL
                for (int j = 0; j < SU; ++j) set_amp(acc[j], 0.0);
                for (int pi = 0; pi < p.passes; ++pi) {
                    if (xo != xprev) {                        // a new input element (predicated loads)
                        QK_UNROLL
                        for (int j = 0; j < SU; ++j) {
                            lo0[j] = *reinterpret_cast<const A*>(row[j] + xo);
                            lo2[j] = *reinterpret_cast<const A*>(row[j] + plane + xo);
                        }
                        xprev = xo;
                    }
                    xp += G;
                    xo = *xp;                                 // next pass (the tables end with one pass of padding steps)
                    select_blocks<A, R, SU, DT>(lo0, lo2, wp, G, acc);
                    wp += wstep;
                }
            };
            if constexpr (SIMPLE) {
                run_row();
                if (k < p.K) {
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j)
                        if (ls + j * SPC < nsamp) amajor_store<A, R>(p, acc[j], out_t, amps_t, o + j * out_stride + k);
                }
            } else {
                for (int b = k; b < p.brows * G_k; b += G_k) {
                    run_row();
                    // UNPREPARE (H on deg) + SUM (H on a) + post-selection deg = a = 0: the sum over the
                    // row's blocks, finished across the G_r lanes with an xor butterfly
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) {
                        for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc[j], shfl_xor_amp(acc[j], m));
                        if (ls + j * SPC < nsamp && r == 0 && b < p.K) amajor_store<A, R>(p, acc[j], out_t, amps_t, o + j * out_stride + b);
                    }
                }
            }
        }
        __syncthreads();                                      // everyone done with cs before the next pre-pass
    }
}

// Direct kernel: layers whose output row b reads ONE input element, x[b N / K] (K a multiple of N: 4x4, 8x8,
// 16x16, 4x8, ... with K a power of two <= 32, one lane per row).  The lane loads that element itself (coalesced:
// the K lanes of a sample read N / K-strided neighbours), runs its CHEB sequence in registers and applies SELECT to
// its row's N (D + 1) blocks - no shared memory, no barriers, no tiles; SU samples per lane share the SELECT
// entries, and the next chunk's inputs are loaded while the current one is evaluated.  An element read by K / N
// rows is evaluated by each of them (K evaluations per sample instead of N).
// GL >= 0: the lanes per sample G = 2^GL are a compile-time constant, so the SELECT loads of a block row are one base
// register + immediate offsets (the run-time-G kernel spent 21 integer instructions per 8 table loads on addresses).
template <class A, typename R, int SU, int NT, int MINB, int DT, int GL>
__global__ void __launch_bounds__(NT, MINB) qkan_block_direct_kernel(const BlockParams p) {
    constexpr int D1 = DT + 1;
    const int G = GL >= 0 ? (1 << GL) : p.G;
    const int SPC = GL >= 0 ? (NT >> GL) : p.SPC;
    const int tid = threadIdx.x;
    const int k = tid & (G - 1);                             // the lane's output row
    const int slot = GL >= 0 ? (tid >> GL) : (tid >> p.g_k_log2);   // sample slot inside the CTA
    if (k >= p.K) return;                                    // no cross-lane step in this kernel: idle lanes leave
    const int nk = (int)(((long long)k * p.N) / p.K);        // the row's input element (ChebyshevStep.py:64, QKANLayer.py:132)
    const bool counts = ((long long)k * p.N) % p.K == 0;     // the first row that reads an element reports its range violation
    const CS<R>* __restrict__ wrow = reinterpret_cast<const CS<R>*>(p.cstab) + k;
    const size_t wstep = (size_t)G * D1;                      // table entries between consecutive a of a lane

    A init[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        init[q].re = (R)p.init[2 * q];
        if constexpr (A::is_complex) init[q].im = (R)p.init[2 * q + 1];
    }
    // a CTA iteration takes a chunk of SU * SPC consecutive samples: lane (slot, k) owns samples slot + j SPC of the chunk.
    // Running pointers / counters (64-bit adds once per chunk, 32-bit offsets inside)
    const long long chunk = (long long)SU * SPC;
    const long long first = (long long)blockIdx.x * chunk + slot;          // the lane's first sample
    const long long adv = (long long)gridDim.x * chunk;
    long long left = p.B - first;                             // samples from the lane's current first one to the end of the batch
    const double* __restrict__ xq = p.x + first * p.N + nk;
    double* __restrict__ oq = p.outs[0] + (p.row0 + first) * p.K + k;
    Cplx<R>* const aq = reinterpret_cast<Cplx<R>*>(p.amps);
    long long eoff = 0;                                       // result elements advanced since the lane's first chunk
    const long long xadv = adv * p.N, oadv = adv * p.K;
    const int xstride = SPC * p.N, ostride = SPC * p.K;
    const bool plain = p.plain != 0;
    unsigned bad = 0;

    double xn[SU];
    QK_UNROLL
    for (int j = 0; j < SU; ++j) xn[j] = ((long long)j * SPC < left) ? xq[j * xstride] : 0.0;
    for (; left > 0; left -= adv, oq += oadv, eoff += oadv) {
        double xv[SU];
        QK_UNROLL
        for (int j = 0; j < SU; ++j) xv[j] = xn[j];
        xq += xadv;
        {   // next chunk's inputs: in flight during this chunk's arithmetic
            const long long nleft = left - adv;
            if (nleft > (long long)(SU - 1) * SPC) {
                QK_UNROLL
                for (int j = 0; j < SU; ++j) xn[j] = xq[j * xstride];
            } else {
                QK_UNROLL
                for (int j = 0; j < SU; ++j) xn[j] = ((long long)j * SPC < nleft) ? xq[j * xstride] : 0.0;
            }
        }
        const bool full = left > (long long)(SU - 1) * SPC;   // every sample of the lane's share exists
        A lo0[SU], lo2[SU], acc[SU];
        // range count (the reference prints a warning, ChebyshevStep.py:46-49) and clip (:52): only inputs with |v| >= 1
        // (or NaN) can need either
        bool edge = false;
        QK_UNROLL
        for (int j = 0; j < SU; ++j) edge |= at_or_beyond_unit(xv[j]);
        if (edge) {
            QK_UNROLL
            for (int j = 0; j < SU; ++j) {
                const double v = xv[j];
                if (!(fabs(v) <= 1.0 + 1e-8) && counts && (full || (long long)j * SPC < left)) ++bad;
                xv[j] = (double)clip_unit<double>(v);
            }
        }
        QK_UNROLL
        for (int j = 0; j < SU; ++j) {
            cheb_element<A, R, DT>(init, (R)xv[j], lo0[j], lo2[j]);      // CHEB sequence of the element
            set_amp(acc[j], 0.0);
        }
        // SELECT on the row's N (D + 1) blocks; UNPREPARE + SUM + post-selection is the lane's running sum
        const CS<R>* wp = wrow;
        for (int a = 0; a < p.N; ++a, wp += wstep) select_blocks<A, R, SU, DT>(lo0, lo2, wp, G, acc);
        if (plain && full) {
            QK_UNROLL
            for (int j = 0; j < SU; ++j) oq[j * ostride] = (double)acc[j].re * p.out_scale;
        } else {
            QK_UNROLL
            for (int j = 0; j < SU; ++j)
                if ((long long)j * SPC < left) {
                    store_result(p, (p.row0 + first) * p.K + k + eoff + j * ostride, (double)acc[j].re * p.out_scale);
                    if (aq) {                                 // amplitudes are local: no row offset
                        Cplx<R> z;
                        z.re = (R)((double)acc[j].re * p.amp_scale);
                        if constexpr (A::is_complex) z.im = (R)((double)acc[j].im * p.amp_scale);
                        else z.im = R(0);
                        aq[first * p.K + k + eoff + j * ostride] = z;
                    }
                }
        }
    }
    if (bad) atomicAdd(p.oor, (unsigned long long)bad);
}

// Element-owner kernel (see ElemLayout above): no shared memory, SU samples per lane, x read straight from global memory
// (the G_r lanes of a row read G_r consecutive inputs), the next element's inputs in flight during the current element's
// arithmetic.  An input shared by two rows (window boundaries) is evaluated by both.
// GL >= 0: rows one after the other (G_k = 1) with G = G_r = 2^GL lanes per row as compile-time constants (immediate
// offsets for the SELECT loads, unrolled butterfly).
template <class A, typename R, int SU, int NT, int MINB, int DT, int GL>
__global__ void __launch_bounds__(NT, MINB) qkan_block_elem_kernel(const BlockParams p) {
    constexpr int D1 = DT + 1;
    const int G = GL >= 0 ? (1 << GL) : p.G;
    const int G_r = GL >= 0 ? (1 << GL) : p.G_r;
    const int SPC = GL >= 0 ? (NT >> GL) : p.SPC;
    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int r = g & (G_r - 1);
    const int k = GL >= 0 ? 0 : (g >> p.g_r_log2);
    const int slot = GL >= 0 ? (tid >> GL) : (tid >> (p.g_r_log2 + p.g_k_log2));
    const CS<R>* __restrict__ we = reinterpret_cast<const CS<R>*>(p.cstab);
    const int* __restrict__ xe = p.xotab;

    A init[4];
    QK_UNROLL
    for (int q = 0; q < 4; ++q) {
        init[q].re = (R)p.init[2 * q];
        if constexpr (A::is_complex) init[q].im = (R)p.init[2 * q + 1];
    }
    const long long chunk = (long long)SU * SPC;
    const long long n_chunks = (p.B + chunk - 1) / chunk;
    const size_t jstep = (size_t)D1 * G;                      // table entries between consecutive j of a lane
    unsigned bad = 0;

    // Loop order.  chunk-outer (p.window == 0): a chunk of samples walks all rows, i.e. the whole SELECT table, before the CTA
    // moves on - fine while the table fits L1.  row-outer (p.window != 0, wide layers): all the CTA's chunks are taken through
    // ONE row before the next row starts, so the row's slice of the table (passes K (D + 1) G entries; N784 K10 D5: 77 KB of
    // the 752 KB) stays in L1 across the chunks of both resident CTAs instead of streaming from L2 once per chunk (the wait
    // for those loads was 39 % of the warp samples at C4, profiles/r02F_ncu_c4.txt).  Every (sample, row) is still evaluated
    // once, by the same code, and x is still read once: a row reads only its own window of the input row.
    const bool row_outer = p.window != 0;
    const int n_ob = row_outer ? p.brows : 1, rows_per = row_outer ? 1 : p.brows;
    for (int ob = 0; ob < n_ob; ++ob)
    for (long long c = blockIdx.x; c < n_chunks; c += gridDim.x) {
        const long long s0 = c * chunk + slot;                // the lane's samples: s0 + j SPC
        const long long left = p.B - s0;                      // <= 0: the whole lane is beyond the batch (still runs: the butterfly needs every lane)
        const double* __restrict__ xs = p.x + (left > 0 ? s0 : 0) * p.N;
        const int xstride = SPC * p.N;
        const int bi0 = ob * rows_per;
        const int* xp = xe + g + (size_t)bi0 * p.passes * G;
        const CS<R>* wp = we + g + (size_t)bi0 * p.passes * p.K * jstep;
        // the first element's inputs
        int en = *xp;
        double xn[SU];
        QK_UNROLL
        for (int j = 0; j < SU; ++j) xn[j] = ((long long)j * SPC < left) ? xs[j * xstride + (en < 0 ? 0 : (en & 0xFFFFF))] : 0.0;
        for (int bi = bi0; bi < bi0 + rows_per; ++bi) {
            A acc[SU];
            QK_UNROLL
            for (int j = 0; j < SU; ++j) set_amp(acc[j], 0.0);
            for (int pi = 0; pi < p.passes; ++pi) {
                const int e_cur = en;
                double xv[SU];
                QK_UNROLL
                for (int j = 0; j < SU; ++j) xv[j] = xn[j];
                xp += G;
                en = *xp;                                     // next element (the table ends with one pass of padding steps)
                {
                    const int off = en < 0 ? 0 : (en & 0xFFFFF);
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) xn[j] = ((long long)j * SPC < left) ? xs[j * xstride + off] : 0.0;
                }
                A lo0[SU], lo2[SU];
                const bool counts = e_cur >= 0 && (e_cur & (1 << 30));
                // range count (ChebyshevStep.py:46-49; once per input) and clip (:52): only for |v| >= 1 or NaN
                bool edge = false;
                QK_UNROLL
                for (int j = 0; j < SU; ++j) edge |= at_or_beyond_unit(xv[j]);
                if (edge) {
                    QK_UNROLL
                    for (int j = 0; j < SU; ++j) {
                        const double v = xv[j];
                        if (counts && (long long)j * SPC < left && !(fabs(v) <= 1.0 + 1e-8)) ++bad;
                        xv[j] = (double)clip_unit<double>(v);
                    }
                }
                QK_UNROLL
                for (int j = 0; j < SU; ++j) cheb_element<A, R, DT>(init, (R)xv[j], lo0[j], lo2[j]);   // CHEB sequence of the element
                // SELECT on the element's K (D + 1) blocks (padding entries rotate by pi: they add exactly 0)
                for (int j = 0; j < p.K; ++j, wp += jstep) select_blocks<A, R, SU, DT>(lo0, lo2, wp, G, acc);
            }
            // UNPREPARE + SUM + post-selection: the sum over the row's blocks, finished across the G_r lanes
            const int b = GL >= 0 ? bi : ((bi << p.g_k_log2) + k);
            QK_UNROLL
            for (int j = 0; j < SU; ++j) {
                QK_UNROLL
                for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc[j], shfl_xor_amp(acc[j], m));
                if (r == 0 && b < p.K && (long long)j * SPC < left) {
                    const long long o = (s0 + (long long)j * SPC) * p.K + b;
                    store_result(p, p.row0 * p.K + o, (double)acc[j].re * p.out_scale);
                    if (p.amps) {
                        Cplx<R> z;
                        z.re = (R)((double)acc[j].re * p.amp_scale);
                        if constexpr (A::is_complex) z.im = (R)((double)acc[j].im * p.amp_scale);
                        else z.im = R(0);
                        reinterpret_cast<Cplx<R>*>(p.amps)[o] = z;
                    }
                }
            }
        }
    }
    if (bad) atomicAdd(p.oor, (unsigned long long)bad);
}

template <typename R>
__global__ void qkan_prepare_elem_tables_kernel(const double* W, int N, int K, int D, int passes, int brows, int g_r_log2, int g_k_log2,
                                                long long steps_total, CS<R>* we, int* xe, unsigned long long* bad_weights) {
    const long long step = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long nw = (long long)N * K * (D + 1);
    unsigned bad = 0;
    for (long long i = step; i < nw; i += (long long)gridDim.x * blockDim.x)
        if (!(fabs(W[i]) <= 1.0)) ++bad;                      // MulStep.py:36-37, each weight once
    if (bad) atomicAdd(bad_weights, (unsigned long long)bad);
    if (step >= steps_total) return;
    fill_elem_step<R>(step, W, N, K, D, passes, brows, g_r_log2, g_k_log2, we, xe);
}

template <class A, typename R, int SU, int NT, int MINB, int DT, int GL>
cudaError_t launch_elem_impl(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    auto kern = qkan_block_elem_kernel<A, R, SU, NT, MINB, DT, GL>;
    BlockParams p = p0;
    const int SPC = NT / G;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long chunk = (long long)SU * SPC;
    const long long n_chunks = (p.B + chunk - 1) / chunk;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_chunks) grid = n_chunks;
    if (grid < 1) grid = 1;
    p.sub = SU; p.tma_ok = 0; p.direct_x = 1; p.s_tot = -1;
    p.G = G; p.G_r = 1 << p.g_r_log2; p.G_k = 1 << p.g_k_log2;
    p.SPC = SPC; p.tile = (int)chunk;
    p.row_bytes = 0; p.plane_bytes = 0;
    // row-outer order when the SELECT table does not fit L1 but one row's slice does (see the kernel)
    {
        const size_t entry = sizeof(CS<R>);
        const size_t row_slice = (size_t)p.passes * p.K * (DT + 1) * G * entry, table = row_slice * p.brows;
        int ro = (table > 96 * 1024 && row_slice <= 128 * 1024 && n_chunks >= 2 * grid) ? 1 : 0;
        if (const char* e = getenv("QKAN_ELEM_ROW_OUTER")) ro = atoi(e) != 0;          // A/B aid
        p.window = ro;
    }
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = 0;
    kern<<<(unsigned)grid, NT, 0, stream>>>(p);
    return cudaGetLastError();
}
// 16 lanes per row (every row wider than ~14 inputs: plan_elem_layout) is compiled in; other splits run the run-time-G kernel
template <class A, typename R, int SU, int NT, int MINB, int DT>
cudaError_t launch_elem(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (p0.D != DT) return cudaErrorInvalidValue;
    static const bool generic_only = getenv("QKAN_ELEM_RUNTIME_G") != nullptr;        // A/B aid
    if (!generic_only && p0.g_k_log2 == 0 && p0.g_r_log2 == 4 && G == 16)
        return launch_elem_impl<A, R, SU, NT, MINB, DT, 4>(p0, G, sm_count, stream, grid_out, smem_out);
    return launch_elem_impl<A, R, SU, NT, MINB, DT, -1>(p0, G, sm_count, stream, grid_out, smem_out);
}
template <class A, typename R, int SU, int NT, int MINB, int DT>
BlockKernelInfo make_elem_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = 0; k.U = 1; k.SU = SU; k.NT = NT; k.MINB = MINB; k.DT = DT; k.is_default = is_default;
    k.tan = 1; k.window = 0; k.amajor = 1; k.direct = 0; k.elem = 1; k.amp_bytes = (int)sizeof(A);
    k.launch = &launch_elem<A, R, SU, NT, MINB, DT>;
    return k;
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
    p.row_bytes = amajor_row_amps(p.N + 1, G, (int)sizeof(A)) * (int)sizeof(A);
    p.plane_bytes = (p.N + 1) * (int)sizeof(A);
    if (const char* e = getenv("QKAN_BLOCK_ROW_AMPS")) {       // tuning aid
        if (atoi(e) >= 2 * (p.N + 1)) p.row_bytes = atoi(e) * (int)sizeof(A);
    }
    // wide rows: staging the raw x twice more than doubles the shared memory per sample and would
    // halve the resident warps; the per-tile compute is long, so the pre-pass reads global memory directly
    p.direct_x = amajor_direct_x(SPC, p.N) ? 1 : 0;
    auto smem_for = [&](int sub) { return amajor_smem_bytes(p.N, SPC, p.row_bytes, SU, sub); };
    int sub = (int)(8192 / ((size_t)SPC * p.N * 8));          // about 8 KiB of x per tile ...
    const int sub_cs = (int)(49152 / ((size_t)SPC * p.row_bytes));   // ... and at most 48 KiB of block amplitudes
    if (sub > sub_cs) sub = sub_cs;
    if (const char* e = getenv("QKAN_BLOCK_SUB")) sub = atoi(e);   // tuning aid
    if (sub > 32) sub = 32;
    // a lane takes SU sub-iterations at a time: a tile of an odd number of them would leave a sample slot idle
    auto round_su = [](int v) { v -= v % SU; return v < SU ? SU : v; };
    sub = round_su(sub);
    while (sub > SU && smem_for(sub) > AMAJOR_SMEM_CAP) sub = round_su(sub - SU);
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
    k.tan = 1; k.window = 0; k.amajor = 1; k.direct = 0; k.elem = 0; k.amp_bytes = (int)sizeof(A);
    k.launch = &launch_amajor<A, R, SU, NT, MINB, DT>;
    return k;
}

template <class A, typename R, int SU, int NT, int MINB, int DT, int GL>
cudaError_t launch_direct_impl(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    auto kern = qkan_block_direct_kernel<A, R, SU, NT, MINB, DT, GL>;
    BlockParams p = p0;
    const int SPC = NT / G;
    int per_sm = 0;
    cudaError_t e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, 0);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long chunk = (long long)SU * SPC;
    const long long n_chunks = (p.B + chunk - 1) / chunk;
    long long grid = (long long)sm_count * per_sm;
    if (grid > n_chunks) grid = n_chunks;
    if (grid < 1) grid = 1;
    p.sub = SU; p.tma_ok = 0; p.direct_x = 1; p.s_tot = -1;
    p.G = G; p.G_r = 1; p.G_k = G;
    p.SPC = SPC; p.tile = (int)chunk;
    p.row_bytes = 0; p.plane_bytes = 0;
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = 0;
    kern<<<(unsigned)grid, NT, 0, stream>>>(p);
    return cudaGetLastError();
}
// G = 4, 8, 16 lanes per sample (K of the BASELINE layers) are compiled in; any other power of two runs the run-time-G kernel
template <class A, typename R, int SU, int NT, int MINB, int DT>
cudaError_t launch_direct(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (p0.D != DT || p0.g_r_log2 != 0 || p0.brows != 1 || p0.K % p0.N != 0 || G != (1 << p0.g_k_log2)) return cudaErrorInvalidValue;
    static const bool generic_only = getenv("QKAN_DIRECT_RUNTIME_G") != nullptr;      // A/B aid
    if (!generic_only) {
        switch (p0.g_k_log2) {
            case 2: return launch_direct_impl<A, R, SU, NT, MINB, DT, 2>(p0, G, sm_count, stream, grid_out, smem_out);
            case 3: return launch_direct_impl<A, R, SU, NT, MINB, DT, 3>(p0, G, sm_count, stream, grid_out, smem_out);
            case 4: return launch_direct_impl<A, R, SU, NT, MINB, DT, 4>(p0, G, sm_count, stream, grid_out, smem_out);
            default: break;
        }
    }
    return launch_direct_impl<A, R, SU, NT, MINB, DT, -1>(p0, G, sm_count, stream, grid_out, smem_out);
}
template <class A, typename R, int SU, int NT, int MINB, int DT>
BlockKernelInfo make_direct_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = 0; k.U = 1; k.SU = SU; k.NT = NT; k.MINB = MINB; k.DT = DT; k.is_default = is_default;
    k.tan = 1; k.window = 0; k.amajor = 1; k.direct = 1; k.elem = 0; k.amp_bytes = (int)sizeof(A);
    k.launch = &launch_direct<A, R, SU, NT, MINB, DT>;
    return k;
}

#endif  // __CUDACC__

}  // namespace qkan
