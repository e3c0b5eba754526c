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

// pick (U, G_r, G_k): maximise slot efficiency, then prefer a register-resident layer
// (one pass, one row per lane), more blocks per lane (ILP), fewer shuffle steps.
// min_g_log2 lets the caller force wide groups (few samples per CTA) for very wide inputs.
inline BlockLayout plan_block_layout(int N, int K, int D, int min_g_log2 = 0, int force_U = 0) {
    const long long rowlen = (long long)N * (D + 1);
    BlockLayout best{};
    double best_score = -1.0;
    const int Us[3] = {4, 2, 1};
    for (int ui = 0; ui < 3; ++ui) {
        const int U = Us[ui];
        if (force_U && U != force_U) continue;
        for (int gr = 0; gr <= 5; ++gr) {
            for (int gk = 0; gr + gk <= 5; ++gk) {
                if (gr + gk < min_g_log2) continue;
                const long long G_r = 1ll << gr, G_k = 1ll << gk;
                const long long passes = (rowlen + G_r * U - 1) / (G_r * U);
                const long long brows = (K + G_k - 1) / G_k;
                const double eff = (double)(rowlen * K) / (double)(G_r * U * passes * G_k * brows);
                const bool resident = passes == 1 && brows == 1;
                // efficiency dominates; the rest only breaks near-ties (within 0.5 %)
                const double score = eff + (resident ? 4e-3 : 0.0) + 1e-3 * U / 4.0 - 1e-4 * gr + 1e-5 * gk;
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

struct BlockParams {
    const double* x;            // [B, N]
    const void* wtab;           // CS<R>[K * rowlen]: (w, sqrt(1-w^2)) of block e = b*rowlen + a*(D+1) + d
    const int* xitab;           // int[K * rowlen]: x index (a + N b) / K  |  d << 20
    double* out;                // [B, K]
    void* amps;                 // optional [B, K] complex
    unsigned long long* oor;
    long long B;
    int N, K, D;
    int rowlen;                 // N * (D + 1) blocks per output row
    int g_r_log2, g_k_log2;     // lanes per row / rows in parallel inside a group
    int passes;                 // ceil(rowlen / (G_r * U))
    int brows;                  // ceil(K / G_k)
    int sub;                    // sub-iterations per x tile
    int tma_ok;
    double out_scale, amp_scale;
};

// one block entry of the tables (host or device)
template <typename R>
QK_HD void fill_block_entry(long long e, const double* W, int N, int K, int D, CS<R>* wtab, int* xitab) {
    const int rowlen = N * (D + 1);
    const int b = (int)(e / rowlen);
    const int i = (int)(e - (long long)b * rowlen);
    const int a = i / (D + 1), d = i - a * (D + 1);
    const int flat = a + N * b;                         // QKANLayer.py:132 (column-major reshape)
    const R w = (R)W[(long long)d * N * K + flat];      // MulStep.py:69
    CS<R> cs;
    cs.c = w;
    cs.s = qk_sqrt((R(1) - w) * (R(1) + w));
    wtab[e] = cs;
    xitab[e] = (flat / K) | (d << 20);                  // ChebyshevStep.py:64 (np.repeat -> i // K)
}

// evolve U blocks and return the sum of their (0,0) amplitudes.  v[u][fx + 2 fw].
template <class A, typename R, int U, int MODE>
QK_HD A evolve_blocks(const R (&cx)[U], const R (&sx)[U], const R (&cw)[U], const R (&sw)[U], const int (&deg)[U], int D) {
    A v[U][4];
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        set_amp(v[u][0], 1.0);          // PREPARE'd, un-normalised
        set_amp(v[u][1], 0.0);
        set_amp(v[u][2], 0.0);
        set_amp(v[u][3], 0.0);
    }
    // CHEB: D applications of the input block-encoding, U and Z U^dagger Z alternating; as real
    // matrices both equal Ry(theta_x), so each application is the same rotation pass
    for (int r = 0; r < D; ++r) {
        QK_UNROLL
        for (int u = 0; u < U; ++u) {
            R c = cx[u], s = sx[u];
            if constexpr (MODE == 1) {                  // paper: term d gets d applications
                const bool on = deg[u] >= r + 1;
                c = on ? c : R(1);
                s = on ? s : R(0);
            }
            rot(v[u][0], v[u][1], c, s);
            rot(v[u][2], v[u][3], c, s);
        }
    }
    // MUL / SELECT on f_w, then read-out of (f_x, f_w) = (0, 0)
    A acc;
    set_amp(acc, 0.0);
    QK_UNROLL
    for (int u = 0; u < U; ++u) {
        rot(v[u][0], v[u][2], cw[u], sw[u]);
        rot(v[u][1], v[u][3], cw[u], sw[u]);
        add_amp(acc, v[u][0]);
    }
    return acc;
}

#if defined(__CUDACC__)
template <class A> __device__ __forceinline__ A shfl_xor_amp(const A& a, int m) {
    A r;
    r.re = __shfl_xor_sync(0xffffffffu, a.re, m);
    if constexpr (A::is_complex) r.im = __shfl_xor_sync(0xffffffffu, a.im, m);
    return r;
}

// RESIDENT: every live block of the layer has its own (lane, u) slot (one pass, one row step), so
// the weight rotations, x offsets and degrees stay in registers for the whole launch and the
// per-sample body is straight-line code.  Otherwise the lane streams its blocks pass by pass.
template <class A, typename R, int U, int MODE, int NT, int MINB, bool RESIDENT>
__global__ void __launch_bounds__(NT, MINB) qkan_block_kernel(const BlockParams p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int G = 1 << (p.g_r_log2 + p.g_k_log2);
    const int G_r = 1 << p.g_r_log2;
    const int G_k = 1 << p.g_k_log2;
    const int SPC = NT / G;                                  // samples in flight per CTA
    const int tile = SPC * p.sub;                            // samples per x tile
    // smem: xs (TMA destination: raw x rows of the next tile) | cs (clip + sqrt of the current tile) | mbar
    const size_t xs_doubles = ((size_t)tile * p.N + 1) & ~(size_t)1;
    double* xs = reinterpret_cast<double*>(smem_raw);
    CS<R>* cs = reinterpret_cast<CS<R>*>(smem_raw + xs_doubles * sizeof(double));
    unsigned long long* mbar =
        reinterpret_cast<unsigned long long*>(smem_raw + xs_doubles * sizeof(double) + (((size_t)tile * p.N * sizeof(CS<R>) + 15) & ~(size_t)15));

    const int tid = threadIdx.x;
    const int g = tid & (G - 1);
    const int r = g & (G_r - 1);
    const int k = g >> p.g_r_log2;
    const int slot = tid >> (p.g_r_log2 + p.g_k_log2);       // sample slot inside the CTA
    const long long n_it = (p.B + tile - 1) / tile;
    const CS<R>* __restrict__ wtab = reinterpret_cast<const CS<R>*>(p.wtab);
    const int* __restrict__ xitab = p.xitab;

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        fence_barrier_init();
    }
    __syncthreads();

    auto tile_bytes = [&](long long it) -> unsigned {
        const long long s0 = it * tile;
        const int ns = (int)((p.B - s0 < tile) ? (p.B - s0) : tile);
        return (unsigned)ns * (unsigned)p.N * 8u;
    };
    auto issue_x = [&](long long it) {
        const unsigned bytes = tile_bytes(it);
        const double* src = p.x + it * tile * p.N;
        if (p.tma_ok && (bytes & 15u) == 0) {
            if (tid == 0) {
                fence_proxy_async();
                mbar_expect_tx(&mbar[0], bytes);
                tma_load_1d(xs, src, bytes, &mbar[0]);
            }
        } else {
            for (int i = tid; i < (int)(bytes >> 3); i += NT) xs[i] = src[i];
        }
    };

    // sample-independent block coefficients of this lane: one load per launch when RESIDENT
    R cw[U], sw[U];
    int xi[U], deg[U];
    bool live[U];
    auto load_items = [&](int b, int i0) {
        QK_UNROLL
        for (int u = 0; u < U; ++u) {
            const int i = i0 + u * G_r;
            live[u] = (b < p.K) && (i < p.rowlen);
            CS<R> w;
            w.c = R(0); w.s = R(1);
            int packed = 0;
            if (live[u]) {
                const long long e = (long long)b * p.rowlen + i;
                w = wtab[e];
                packed = xitab[e];
            }
            cw[u] = w.c; sw[u] = w.s;
            xi[u] = packed & 0xFFFFF;
            deg[u] = packed >> 20;
        }
    };
    if constexpr (RESIDENT) load_items(k, r);

    long long it = blockIdx.x;
    unsigned phase = 0;
    if (it < n_it) issue_x(it);
    __syncthreads();

    const size_t row_stride = (size_t)SPC * p.N;             // cs entries between consecutive sub-iterations
    const long long out_stride = (long long)SPC * p.K;

    for (; it < n_it; it += gridDim.x) {
        if (p.tma_ok && (tile_bytes(it) & 15u) == 0) { mbar_wait(&mbar[0], phase); phase ^= 1; }
        const long long s0 = it * tile;
        const int nsamp = (int)((p.B - s0 < tile) ? (p.B - s0) : tile);

        // pre-pass over the raw inputs of the tile: range count (the reference prints a warning,
        // ChebyshevStep.py:46-49), clip (:52) and the rotation pair cos(theta/2) = x,
        // sin(theta/2) = sqrt(1 - x^2) - no arccos is ever needed
        unsigned bad = 0;
        for (int i = tid; i < nsamp * p.N; i += NT) {
            const double v = xs[i];
            if (!(-1.0 - 1e-8 <= v) || !(v <= 1.0 + 1e-8)) ++bad;
            const R c = clip_unit<R>(v);
            CS<R> e;
            e.c = c;
            e.s = qk_sqrt((R(1) - c) * (R(1) + c));
            cs[i] = e;
        }
        if (bad) atomicAdd(p.oor, (unsigned long long)bad);
        __syncthreads();                                      // cs complete, xs free again
        const long long nxt = it + gridDim.x;
        if (nxt < n_it) issue_x(nxt);                         // overlaps with the compute below

        const int nsub = (nsamp + SPC - 1) / SPC;
        const CS<R>* csrow = cs + (size_t)slot * p.N;
        long long o = (s0 + slot) * p.K;
        int ls = slot;
        for (int si = 0; si < nsub; ++si, csrow += row_stride, o += out_stride, ls += SPC) {
            const bool valid = ls < nsamp;
            const CS<R>* row = valid ? csrow : cs;            // idle slots of a ragged tile read row 0
            if constexpr (RESIDENT) {
                R cx[U], sx[U];
                QK_UNROLL
                for (int u = 0; u < U; ++u) {
                    CS<R> e;
                    e.c = R(0); e.s = R(1);
                    if (live[u]) e = row[xi[u]];
                    cx[u] = e.c; sx[u] = e.s;
                }
                A acc = evolve_blocks<A, R, U, MODE>(cx, sx, cw, sw, deg, p.D);
                // UNPREPARE (H on deg) + SUM (H on a) + post-selection deg = a = 0: the sum over the
                // row's blocks, finished across the G_r lanes with an xor butterfly
                for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc, shfl_xor_amp(acc, m));
                if (valid && r == 0 && k < p.K) {
                    p.out[o + k] = (double)acc.re * p.out_scale;
                    if (p.amps) {
                        Cplx<R> z;
                        z.re = (R)((double)acc.re * p.amp_scale);
                        if constexpr (A::is_complex) z.im = (R)((double)acc.im * p.amp_scale);
                        else z.im = R(0);
                        reinterpret_cast<Cplx<R>*>(p.amps)[o + k] = z;
                    }
                }
            } else {
                for (int b = k; b < p.brows * G_k; b += G_k) {
                    A acc;
                    set_amp(acc, 0.0);
                    int i0 = r;
                    for (int pi = 0; pi < p.passes; ++pi, i0 += U * G_r) {
                        load_items(b, i0);
                        R cx[U], sx[U];
                        QK_UNROLL
                        for (int u = 0; u < U; ++u) {
                            CS<R> e;
                            e.c = R(0); e.s = R(1);
                            if (live[u]) e = row[xi[u]];
                            cx[u] = e.c; sx[u] = e.s;
                        }
                        const A part = evolve_blocks<A, R, U, MODE>(cx, sx, cw, sw, deg, p.D);
                        add_amp(acc, part);
                    }
                    for (int m = G_r >> 1; m >= 1; m >>= 1) add_amp(acc, shfl_xor_amp(acc, m));
                    if (valid && r == 0 && b < p.K) {
                        p.out[o + b] = (double)acc.re * p.out_scale;
                        if (p.amps) {
                            Cplx<R> z;
                            z.re = (R)((double)acc.re * p.amp_scale);
                            if constexpr (A::is_complex) z.im = (R)((double)acc.im * p.amp_scale);
                            else z.im = R(0);
                            reinterpret_cast<Cplx<R>*>(p.amps)[o + b] = z;
                        }
                    }
                }
            }
        }
        __syncthreads();                                      // everyone done with cs before the next pre-pass
    }
}

template <typename R>
__global__ void qkan_prepare_block_tables_kernel(const double* W, int N, int K, int D, CS<R>* wtab, int* xitab,
                                                 unsigned long long* bad_weights) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long E = (long long)N * K * (D + 1);
    if (e >= E) return;
    fill_block_entry<R>(e, W, N, K, D, wtab, xitab);
    const int rowlen = N * (D + 1);
    const int b = (int)(e / rowlen), i = (int)(e % rowlen);
    const int a = i / (D + 1), d = i % (D + 1);
    const double w = W[(long long)d * N * K + a + N * b];
    if (!(fabs(w) <= 1.0)) atomicAdd(bad_weights, 1ull);     // MulStep.py:36-37
}

struct BlockKernelInfo {
    int amp, mode, U, NT, MINB;
    int is_default;
    cudaError_t (*launch)(const BlockParams&, int g, int sm_count, cudaStream_t, int* grid_out, int* smem_out);
};

template <class A, typename R, int U, int MODE, int NT, int MINB, bool RESIDENT>
cudaError_t launch_block_impl(const BlockParams& p0, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    auto kern = qkan_block_kernel<A, R, U, MODE, NT, MINB, RESIDENT>;
    BlockParams p = p0;
    const int SPC = NT / G;
    auto smem_for = [&](int sub) {
        const size_t tile = (size_t)SPC * sub;
        const size_t xs = ((tile * p.N + 1) & ~(size_t)1) * sizeof(double);
        const size_t cs = (tile * p.N * sizeof(CS<R>) + 15) & ~(size_t)15;
        return xs + cs + 16;
    };
    int sub = (int)(8192 / ((size_t)SPC * p.N * 8));          // about 8 KiB of x per tile
    if (sub > 32) sub = 32;
    if (sub < 1) sub = 1;
    if (smem_for(sub) > 200 * 1024) return cudaErrorInvalidConfiguration;
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
    p.tma_ok = ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0 && (((size_t)SPC * sub * p.N * 8) & 15u) == 0) ? 1 : 0;
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = (int)smem_for(sub);
    kern<<<(unsigned)grid, NT, smem_for(sub), stream>>>(p);
    return cudaGetLastError();
}
template <class A, typename R, int U, int MODE, int NT, int MINB>
cudaError_t launch_block(const BlockParams& p, int G, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    if (p.passes == 1 && p.brows == 1)
        return launch_block_impl<A, R, U, MODE, NT, MINB, true>(p, G, sm_count, stream, grid_out, smem_out);
    return launch_block_impl<A, R, U, MODE, NT, MINB, false>(p, G, sm_count, stream, grid_out, smem_out);
}

template <class A> struct AmpId;

template <class A, typename R, int U, int MODE, int NT, int MINB>
BlockKernelInfo make_block_info(int is_default) {
    BlockKernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = MODE; k.U = U; k.NT = NT; k.MINB = MINB; k.is_default = is_default;
    k.launch = &launch_block<A, R, U, MODE, NT, MINB>;
    return k;
}
#endif  // __CUDACC__

}  // namespace qkan
