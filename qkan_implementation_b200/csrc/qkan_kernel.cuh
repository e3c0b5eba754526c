// qkan_kernel.cuh - the fused, persistent sm_100a kernel around qkan_core.cuh.
//
// One CTA = NT threads = SPI thread groups of G = 2^(QT-T) threads; each group owns one
// sample's tile statevector (2^QT amplitudes) in shared memory + registers and runs the
// whole gate list on it (run_stage<0..ns-1>), looping over the sectors of the a / b qubits
// that do not fit the tile.  CTAs are persistent: iteration `it` handles samples
// [it*SPI, it*SPI+SPI); the x rows of the next iteration are staged by a 1-D TMA bulk copy
// (cp.async.bulk + mbarrier, double buffered) while the current one is computed.
#pragma once
#include "qkan_core.cuh"
#include <cuda_runtime.h>

namespace qkan {

struct LaunchParams {
    const double* x;         // [B, N] row-major, device
    const void* wtab;        // CS<R>[Kpad * Npad << L]
    const int* xidx;         // int[Kpad * Npad]
    double* out;             // [B, K]
    void* amps;              // optional: [B, K] complex (2 x R) post-selected amplitudes
    unsigned long long* oor; // count of x entries outside [-1-1e-8, 1+1e-8] (ChebyshevStep.py:46)
    long long B;
    int N, K, D, NA, NB;
    double out_scale;        // 1 / (N (D+1))
    double amp_scale;        // 2^-(m + 2l + n_a)/2
    int tma_ok;              // x is 16-byte aligned and a full tile is a multiple of 16 bytes
    int sub;                 // sub-iterations per x tile (tile = SPI * sub samples)
};

// ------------------------------------------------------------------ PTX bits
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    unsigned done = 0;
    const unsigned addr = smem_u32(bar);
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
// 1-D TMA: global -> shared, completion signalled on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

template <int G, int NT> __device__ __forceinline__ void group_sync(int slot) {
    if constexpr (G <= 32) {
        __syncwarp();
    } else if constexpr (G == NT) {
        __syncthreads();
    } else {
        asm volatile("bar.sync %0, %1;" ::"r"(slot + 1), "r"(G) : "memory");
    }
}

template <class A, typename R, class P, int MODE, bool DIRECT, int G, int NT, int S>
struct StageRunner {
    static __device__ __forceinline__ void run(A* st, unsigned t, const TileArgs<R>& ta, const MuxCoef<R, P>& mc,
                                               A* acc, int slot) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        if constexpr (S < p.ns) {
            run_stage<A, R, P, S, MODE, DIRECT>(st, t, ta, mc, acc);
            if constexpr (S + 1 < p.ns) group_sync<G, NT>(slot);
            StageRunner<A, R, P, MODE, DIRECT, G, NT, S + 1>::run(st, t, ta, mc, acc, slot);
        }
    }
};

template <class A, typename R, int L, int NAT, int NBT, int T, int NT, int MINB, int MODE, int PREP, bool FULL>
struct KernelCfg {
    static constexpr int FW = sizeof(A) == 16 ? 3 : (sizeof(A) == 8 ? 4 : 5);
    using P = Plan<L, NAT, NBT, T, FW, PREP>;
    static constexpr int QT = P::QT;
    static constexpr int ST = 1 << QT;
    static constexpr int G = 1 << (QT - T);
    static_assert(G <= NT, "thread group larger than the CTA");
    static constexpr int SPI = NT / G;                 // samples in flight per CTA
    static_assert(G <= 32 || G == NT || SPI <= 15, "not enough named barriers");
    static constexpr size_t state_bytes = (size_t)SPI * ST * sizeof(A);

    // x tile = SPI * sub samples (sub sub-iterations between two CTA-wide barriers)
    static size_t xs_stride(int N, int sub) { return ((size_t)SPI * sub * N + 1) & ~(size_t)1; }   // doubles, 16-B aligned
    static __host__ __device__ size_t acc_bytes(int NB) { return (((size_t)SPI << NB) * sizeof(A) + 15) & ~(size_t)15; }
    static size_t smem_bytes(int N, int NB, int sub) {
        return state_bytes + acc_bytes(NB) + 2 * xs_stride(N, sub) * sizeof(double) + 2 * sizeof(unsigned long long);
    }
};

template <class A, typename R, int L, int NAT, int NBT, int T, int NT, int MINB, int MODE, int PREP, bool FULL>
__global__ void __launch_bounds__(NT, MINB) qkan_forward_kernel(const LaunchParams p) {
    using C = KernelCfg<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    using P = typename C::P;
    constexpr int G = C::G, SPI = C::SPI, ST = C::ST;
    constexpr P plan = make_plan<L, NAT, NBT, T, C::FW, PREP>();

    extern __shared__ __align__(128) unsigned char smem_raw[];
    A* state = reinterpret_cast<A*>(smem_raw);
    const int Kpad = 1 << p.NB;
    A* accs = reinterpret_cast<A*>(smem_raw + C::state_bytes);
    double* xs = reinterpret_cast<double*>(smem_raw + C::state_bytes + C::acc_bytes(p.NB));
    const int sub = p.sub;
    const int tile = SPI * sub;                              // samples per CTA iteration
    const size_t xstride = ((size_t)tile * p.N + 1) & ~(size_t)1;
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(xs + 2 * xstride);

    const int tid = threadIdx.x;
    const int slot = tid / G;
    const unsigned t = tid % G;
    const long long n_it = (p.B + tile - 1) / tile;
    const int n_ahi = 1 << (p.NA - NAT), n_bhi = 1 << (p.NB - NBT);
    const bool one_sector = (n_ahi == 1 && n_bhi == 1);

    if (tid == 0) {
        mbar_init(&mbar[0], 1);
        mbar_init(&mbar[1], 1);
        fence_barrier_init();
    }
    __syncthreads();

    // stage the x rows of iteration `it` into buffer `b`: one 1-D TMA bulk copy when the tile
    // is 16-byte granular, plain coalesced loads otherwise (ragged tail, odd N)
    auto tile_bytes = [&](long long it) -> unsigned {
        const long long s0 = it * tile;
        const int ns = (int)((p.B - s0 < tile) ? (p.B - s0) : tile);
        return (unsigned)ns * (unsigned)p.N * 8u;
    };
    auto issue_x = [&](long long it, int b) {
        const unsigned bytes = tile_bytes(it);
        const double* src = p.x + it * tile * p.N;
        double* dst = xs + (size_t)b * xstride;
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

    long long it = blockIdx.x;
    int buf = 0;
    unsigned phase0 = 0, phase1 = 0;
    if (it < n_it) issue_x(it, 0);
    __syncthreads();

    A* st = state + (size_t)slot * ST;
    A* acc = accs + ((size_t)slot << p.NB);

    TileArgs<R> ta;
    ta.wtab = reinterpret_cast<const CS<R>*>(p.wtab);
    ta.xidx = p.xidx;
    ta.NA = p.NA;
    ta.D = p.D;
    ta.K = p.K;
    ta.a_hi = 0;
    ta.b_hi = 0;
    ta.xrow = nullptr;
    ta.out_row = nullptr;
    ta.amp_row = nullptr;
    ta.out_scale = p.out_scale;
    ta.amp_scale = p.amp_scale;
    MuxCoef<R, P> mc;
    if (one_sector) load_mux_coefs<R, P>(mc, t, ta);     // weight rotations stay in registers for the whole launch

    for (; it < n_it; it += gridDim.x, buf ^= 1) {
        const long long nxt = it + gridDim.x;
        if (nxt < n_it) issue_x(nxt, buf ^ 1);
        if (p.tma_ok && (tile_bytes(it) & 15u) == 0) {
            if (buf == 0) { mbar_wait(&mbar[0], phase0); phase0 ^= 1; }
            else          { mbar_wait(&mbar[1], phase1); phase1 ^= 1; }
        }
        const long long s0 = it * tile;
        const int nsamp = (int)((p.B - s0 < tile) ? (p.B - s0) : tile);
        const double* xt = xs + (size_t)buf * xstride;

        // range check of the raw inputs (the reference prints a warning and clips)
        unsigned bad = 0;
        for (int i = tid; i < nsamp * p.N; i += NT) {
            const double v = xt[i];
            if (!(-1.0 - 1e-8 <= v) || !(v <= 1.0 + 1e-8)) ++bad;
        }
        if (bad) atomicAdd(p.oor, (unsigned long long)bad);

        // sub-iterations: SPI samples at a time, only group-level synchronisation inside
        const int nsub = (nsamp + SPI - 1) / SPI;
        for (int si = 0; si < nsub; ++si) {
            const int ls = si * SPI + slot;                   // sample within the tile
            const bool valid = ls < nsamp;
            ta.xrow = xt + (size_t)(valid ? ls : 0) * p.N;
            if constexpr (FULL) {
                // whole register on chip: read-out goes from registers to global memory
                const long long o = (s0 + ls) * p.K;
                ta.out_row = valid ? p.out + o : nullptr;
                ta.amp_row = (valid && p.amps) ? (void*)(reinterpret_cast<Cplx<R>*>(p.amps) + o) : nullptr;
                if constexpr (PREP == 0) {
                    for (int i = (int)t; i < ST; i += G) set_amp(st[i], i == 0 ? 1.0 : 0.0);
                    group_sync<G, NT>(slot);
                }
                StageRunner<A, R, P, MODE, true, G, NT, 0>::run(st, t, ta, mc, acc, slot);
                // the next sample's first store must not overtake this sample's last loads
                if constexpr (plan.ns > 1) group_sync<G, NT>(slot);
                continue;
            }
            for (int b = (int)t; b < Kpad; b += G) set_amp(acc[b], 0.0);
            group_sync<G, NT>(slot);
            for (int bh = 0; bh < n_bhi; ++bh) {
                for (int ah = 0; ah < n_ahi; ++ah) {
                    if (sector_is_padding(ah, bh, NAT, NBT, p.N, p.K)) continue;
                    ta.a_hi = ah;
                    ta.b_hi = bh;
                    if constexpr (PREP == 0) {
                        // |0...0> in shared memory; the initial Hadamards run as ordinary passes
                        for (int i = (int)t; i < ST; i += G) set_amp(st[i], i == 0 ? 1.0 : 0.0);
                        group_sync<G, NT>(slot);
                    }
                    load_mux_coefs<R, P>(mc, t, ta);
                    StageRunner<A, R, P, MODE, false, G, NT, 0>::run(st, t, ta, mc, acc, slot);
                    if constexpr (plan.ns > 1) group_sync<G, NT>(slot);
                }
            }
            if constexpr (plan.ns == 1) group_sync<G, NT>(slot);
            if (valid) {
                for (int b = (int)t; b < p.K; b += G) {
                    const A a = acc[b];
                    const long long o = (s0 + ls) * p.K + b;
                    p.out[o] = (double)a.re * p.out_scale;
                    if (p.amps) {
                        R im = 0;
                        if constexpr (A::is_complex) im = a.im;
                        Cplx<R> z;
                        z.re = (R)((double)a.re * p.amp_scale);
                        z.im = (R)((double)im * p.amp_scale);
                        reinterpret_cast<Cplx<R>*>(p.amps)[o] = z;
                    }
                }
            }
            group_sync<G, NT>(slot);      // acc is re-zeroed by the next sub-iteration
        }
        __syncthreads();                  // everyone is done with xs[buf] before it is refilled
    }
}

// ------------------------------------------------------------------ launcher
struct KernelInfo {
    int amp;      // 0 = complex128, 1 = complex64, 2 = real64 (real-only representation)
    int mode;     // 0 = compat, 1 = paper
    int prep;     // 1 = closed-form state preparation, 0 = initial Hadamards executed as gates
    int prio;     // selection priority among kernels that fit a shape
    int variant;  // tuning variant id (env QKAN_VARIANT=<id> prefers it); 0 = default
    int full;     // 1: only for shapes whose whole register is the tile (direct read-out, no sector loop)
    int L, NAT, NBT, T, NT;
    int stages;
    int spi;
    cudaError_t (*launch)(const LaunchParams&, int sm_count, cudaStream_t, int* grid_out, int* smem_out);
    const void* func;
};

template <class A, typename R, int L, int NAT, int NBT, int T, int NT, int MINB, int MODE, int PREP, bool FULL>
cudaError_t launch_forward(const LaunchParams& p, int sm_count, cudaStream_t stream, int* grid_out, int* smem_out) {
    using C = KernelCfg<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    auto kern = qkan_forward_kernel<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    // x tile: about 4 KiB per buffer, but keep >= 4 tiles per resident CTA for load balance
    int sub = (int)(4096 / ((size_t)C::SPI * p.N * 8));
    if (sub > 32) sub = 32;
    if (sub < 1) sub = 1;
    size_t smem = C::smem_bytes(p.N, p.NB, sub);
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT, smem);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorLaunchOutOfResources;
    const long long resident = (long long)sm_count * per_sm;
    while (sub > 1 && (p.B + (long long)C::SPI * sub - 1) / ((long long)C::SPI * sub) < 4 * resident) sub >>= 1;
    smem = C::smem_bytes(p.N, p.NB, sub);
    const long long n_it = (p.B + (long long)C::SPI * sub - 1) / ((long long)C::SPI * sub);
    long long grid = resident;
    if (grid > n_it) grid = n_it;
    if (grid < 1) grid = 1;
    LaunchParams q = p;
    q.sub = sub;
    q.tma_ok = ((reinterpret_cast<uintptr_t>(p.x) & 15u) == 0 && (((size_t)C::SPI * sub * p.N * 8) & 15u) == 0) ? 1 : 0;
    if (grid_out) *grid_out = (int)grid;
    if (smem_out) *smem_out = (int)smem;
    kern<<<(unsigned)grid, NT, smem, stream>>>(q);
    return cudaGetLastError();
}

template <class A> struct AmpId;
template <> struct AmpId<Cplx<double>> { static constexpr int v = 0; };
template <> struct AmpId<Cplx<float>> { static constexpr int v = 1; };
template <> struct AmpId<Real<double>> { static constexpr int v = 2; };

template <class A, typename R, int L, int NAT, int NBT, int T, int NT, int MINB, int MODE, int PREP, bool FULL>
KernelInfo make_info() {
    using C = KernelCfg<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    constexpr auto plan = make_plan<L, NAT, NBT, T, C::FW, PREP>();
    KernelInfo k;
    k.amp = AmpId<A>::v;
    k.mode = MODE;
    k.prep = PREP;
    k.prio = 0;
    k.variant = 0;
    k.full = FULL ? 1 : 0;
    k.L = L; k.NAT = NAT; k.NBT = NBT; k.T = T; k.NT = NT;
    k.stages = plan.ns;
    k.spi = C::SPI;
    k.launch = &launch_forward<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    k.func = (const void*)qkan_forward_kernel<A, R, L, NAT, NBT, T, NT, MINB, MODE, PREP, FULL>;
    return k;
}

// table preparation (one thread per padded (b, a) entry)
template <typename R>
__global__ void qkan_prepare_tables_kernel(const double* W, int N, int K, int D, int NA, int NB, int L,
                                           CS<R>* wtab, int* xidx, unsigned long long* bad_weights) {
    const unsigned ab = blockIdx.x * blockDim.x + threadIdx.x;
    if (ab >= (1u << (NA + NB))) return;
    fill_tables_entry<R>(ab, W, N, K, D, NA, L, wtab, xidx);
    // |w| <= 1 is required for the rotation to exist (MulStep.py:36-37)
    const int a = (int)(ab & ((1u << NA) - 1u)), b = (int)(ab >> NA);
    if (a < N && b < K) {
        unsigned bad = 0;
        for (int d = 0; d <= D; ++d) {
            const double w = W[(long long)d * N * K + a + N * b];
            if (!(fabs(w) <= 1.0)) ++bad;
        }
        if (bad) atomicAdd(bad_weights, (unsigned long long)bad);
    }
}

}  // namespace qkan
