// Host emulation of the CUDA kernel's stage machine (TEST TOOL, not product code).
// Compiles qkan_core.cuh with g++ and runs the thread groups sequentially: inside a
// stage every thread reads and writes only its own amplitudes, and stages are separated
// by barriers on the GPU, so sequential in-place execution is equivalent.  Lets the CPU
// test-suite check the compile-time plans, swizzle and read-out against the oracle.
#include "../../qkan_implementation_b200/csrc/qkan_core.cuh"
#include "../../qkan_implementation_b200/csrc/qkan_block.cuh"
#include "../../qkan_implementation_b200/csrc/qkan_amajor.cuh"
#include <vector>
#include <cstring>
#include <cstdio>

using namespace qkan;

template <class A, typename R, class P, int MODE, bool DIRECT, int S>
struct StageLoop {
    static void run(A* state, const TileArgs<R>& ta, A* acc) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        if constexpr (S < p.ns) {
            for (unsigned t = 0; t < (1u << P::NL); ++t) {
                MuxCoef<R, P> mc;
                load_mux_coefs<R, P>(mc, t, ta);
                run_stage<A, R, P, S, MODE, DIRECT>(state, t, ta, mc, acc);
            }
            StageLoop<A, R, P, MODE, DIRECT, S + 1>::run(state, ta, acc);
        }
    }
};

template <class A, typename R, int L, int NAT, int NBT, int T, int MODE, int PREP>
int emu_run(const double* x, const double* W, long long B, int N, int K, int D, double* out, double* amps) {
    using P = Plan<L, NAT, NBT, T, (sizeof(A) == 16 ? 3 : (sizeof(A) == 8 ? 4 : 5)), PREP>;
    int NA = 0, NB = 0, LL = 0;
    while ((1 << NA) < N) ++NA;
    while ((1 << NB) < K) ++NB;
    while ((1 << LL) < D + 1) ++LL;
    if (LL != L || NAT > NA || NBT > NB) return -1;
    const int Np = 1 << NA, Kp = 1 << NB;
    std::vector<CS<R>> wtab((size_t)Np * Kp << L);
    std::vector<int> xidx((size_t)Np * Kp);
    for (unsigned ab = 0; ab < (unsigned)(Np * Kp); ++ab) fill_tables_entry<R>(ab, W, N, K, D, NA, L, wtab.data(), xidx.data());
    std::vector<A> state((size_t)1 << P::QT);
    std::vector<A> acc(Kp);
    const int n_ahi = 1 << (NA - NAT), n_bhi = 1 << (NB - NBT);
    const double amp_scale = std::pow(2.0, -0.5 * (NA + NB + 2 * L + NA));
    bool direct = false;
    for (long long s = 0; s < B; ++s) {
        for (auto& a : acc) set_amp(a, 0.0);
        for (int bh = 0; bh < n_bhi; ++bh)
            for (int ah = 0; ah < n_ahi; ++ah) {
                if (sector_is_padding(ah, bh, NAT, NBT, N, K)) continue;
                TileArgs<R> ta{wtab.data(), xidx.data(), x + s * N, NA, ah, bh, D, K, nullptr, nullptr,
                               1.0 / ((double)N * (D + 1)), amp_scale};
                if (PREP == 0) for (size_t i = 0; i < state.size(); ++i) set_amp(state[i], i == 0 ? 1.0 : 0.0);
                if (NAT == NA && NBT == NB) {      // whole register in the tile: direct read-out path
                    std::vector<Cplx<R>> arow(K);
                    ta.out_row = out + s * K;
                    ta.amp_row = amps ? arow.data() : nullptr;
                    StageLoop<A, R, P, MODE, true, 0>::run(state.data(), ta, acc.data());
                    if (amps) for (int b = 0; b < K; ++b) { amps[2 * (s * K + b)] = arow[b].re; amps[2 * (s * K + b) + 1] = arow[b].im; }
                    direct = true;
                } else {
                    StageLoop<A, R, P, MODE, false, 0>::run(state.data(), ta, acc.data());
                }
            }
        if (!direct) for (int b = 0; b < K; ++b) {
            out[s * K + b] = (double)acc[b].re / ((double)N * (D + 1));
            if (amps) {
                amps[2 * (s * K + b)] = (double)acc[b].re * amp_scale;
                if constexpr (A::is_complex) amps[2 * (s * K + b) + 1] = (double)acc[b].im * amp_scale;
                else amps[2 * (s * K + b) + 1] = 0.0;
            }
        }
    }
    return 0;
}

#define CASE(id, A, R, L, NAT, NBT, T, MODE) \
    case id: return emu_run<A<R>, R, L, NAT, NBT, T, MODE, 1>(x, W, B, N, K, D, out, amps); \
    case 100 + id: return emu_run<A<R>, R, L, NAT, NBT, T, MODE, 0>(x, W, B, N, K, D, out, amps);

extern "C" int qkan_emu_forward(int cfg, const double* x, const double* W, long long B, int N, int K, int D,
                                double* out, double* amps) {
    switch (cfg) {
        CASE(0, Cplx, double, 2, 2, 2, 4, 0)    // C2: N4 K4 D3 full tile
        CASE(1, Cplx, double, 2, 2, 2, 3, 0)    // same, T=3
        CASE(2, Cplx, double, 2, 2, 2, 5, 0)    // same, T=5
        CASE(3, Cplx, double, 2, 2, 0, 4, 0)    // b as sectors
        CASE(4, Cplx, double, 2, 1, 0, 3, 0)    // a_hi and b sectors
        CASE(5, Cplx, float, 2, 2, 2, 4, 0)     // complex64
        CASE(6, Real, double, 2, 2, 2, 4, 0)    // real-only
        CASE(7, Cplx, double, 2, 2, 2, 4, 1)    // paper mode
        CASE(8, Cplx, double, 4, 4, 4, 5, 0)    // C3 full (14 qubits)
        CASE(9, Cplx, double, 4, 4, 2, 4, 0)    // C3 with 2 b qubits as sectors
        CASE(10, Cplx, double, 5, 3, 3, 5, 0)   // C5 D=16
        CASE(11, Cplx, double, 1, 3, 3, 4, 0)   // C5 D=1
        CASE(12, Cplx, double, 3, 2, 1, 4, 0)   // N3 K2 D4
        CASE(13, Cplx, double, 1, 3, 2, 4, 0)   // N5 K3 D1
        CASE(14, Cplx, double, 0, 0, 0, 2, 0)   // N1 K1 D0
        CASE(15, Cplx, double, 3, 7, 0, 4, 0)   // C4: N784 K10 D5, a_hi = 3, b = 4 sector qubits
        CASE(16, Cplx, double, 3, 3, 3, 4, 0)   // N8 K8 D5
        CASE(17, Cplx, double, 3, 3, 3, 4, 1)   // N8 K8 D5 paper
        CASE(18, Cplx, double, 1, 1, 1, 2, 0)   // N2 K2 D1
        CASE(19, Cplx, double, 4, 2, 2, 4, 0)   // N4 K4 D10
    }
    return -2;
}

// ---- block engine: lanes of a group run sequentially, the xor butterfly is replayed on arrays
template <class A, typename R, int U, int DT> struct TanRun {
    static void go(int D, const A (&init)[4], const R (&t)[U], const R (&al)[U], const R (&be)[U], const R (&cw)[U],
                   const R (&sw)[U], A& acc) {
        if (D == DT) evolve_blocks_tan<A, R, U, DT>(init, t, al, be, cw, sw, acc);
        else TanRun<A, R, U, DT - 1>::go(D, init, t, al, be, cw, sw, acc);
    }
};
template <class A, typename R, int U> struct TanRun<A, R, U, 0> {
    static void go(int, const A (&)[4], const R (&)[U], const R (&)[U], const R (&)[U], const R (&)[U], const R (&)[U], A&) {}
};

// tan != 0: the scaled-rotation form of the degree-specialised kernels (compat mode, 1 <= D <= 16);
// tan == 2: the window kernel's tables and per-row-step input windows (U = 1 layouts only)
template <class A, typename R, int U, int MODE>
int emu_block(const double* x, const double* W, long long B, int N, int K, int D, int min_g, int tan, double* out, double* amps) {
    if (tan) return -4;              // the scaled-rotation kernels are a-major now: emu_amajor below
    const BlockLayout lay = plan_block_layout(N, K, D, min_g);
    if (lay.U != U) return -3;
    const int G_r = 1 << lay.g_r_log2, G_k = 1 << lay.g_k_log2, G = G_r * G_k;
    const long long slots = ((long long)lay.brows * lay.passes + 1) * U * G;
    const int window = tan == 2 ? block_window_max(N, K, lay.g_k_log2, lay.brows) : 0;
    std::vector<CS<R>> cstab(slots);
    std::vector<int> xotab(slots);
    for (long long e = 0; e < slots; ++e)
        fill_block_slot<R>(e, W, N, K, D, U, lay.passes, lay.g_r_log2, lay.g_k_log2, MODE, cstab.data(), xotab.data(),
                           tan ? (int)sizeof(TanEntry<R>) : (int)sizeof(CS<R>), window);
    const int NP = tan ? (tan_row_words(N, G, (int)sizeof(R)) + 2) / 3 : N + 1;   // emulation rows: whole triples
    int NA = 0, NB = 0, L = 0;
    while ((1 << NA) < N) ++NA;
    while ((1 << NB) < K) ++NB;
    while ((1 << L) < D + 1) ++L;
    const double amp_scale = std::pow(2.0, -0.5 * (NA + NB + 2 * L + NA));
    for (long long s = 0; s < B; ++s) {
        std::vector<CS<R>> cs(N + 1);
        for (int n = 0; n < N; ++n) { R c = clip_unit<R>(x[s * N + n]); cs[n].c = c; cs[n].s = qk_sqrt((R(1) - c) * (R(1) + c)); }
        cs[N].c = 0; cs[N].s = 1;
        std::vector<TanEntry<R>> cst((size_t)NP);            // (t, alpha, beta) triples
        if (tan)
            for (int n = 0; n <= N; ++n) cst[n] = tan_entry<R>(n < N ? cs[n].c : R(0), D);
        for (int bi = 0; bi < lay.brows; ++bi) {
            if (window) {                                    // the row step's window of the input row, dummy at index `window`
                int lo, len;
                block_window(N, K, lay.g_k_log2, bi, &lo, &len);
                if (len > window) return -6;
                cst.assign((size_t)window + 1, TanEntry<R>{R(7), R(7), R(7)});   // stale entries must never be read
                for (int j = 0; j < len; ++j) cst[j] = tan_entry<R>(cs[lo + j].c, D);
                cst[window] = tan_entry<R>(R(0), D);
            }
            for (int k = 0; k < G_k; ++k) {
                const int b = bi * G_k + k;
                std::vector<A> acc(G_r);
                for (int r = 0; r < G_r; ++r) {
                    set_amp(acc[r], 0.0);
                    for (int pi = 0; pi < lay.passes; ++pi) {
                        R cx[U], sx[U], cw[U], sw[U], tx[U], ax[U], bx[U];
                        int deg[U];
                        const int g = k * G_r + r;
                        for (int u = 0; u < U; ++u) {
                            const size_t sl = (((size_t)bi * lay.passes + pi) * U + u) * G + g;
                            cw[u] = cstab[sl].c; sw[u] = cstab[sl].s;
                            int xo = xotab[sl];
                            deg[u] = MODE == 1 ? (xo >> 24) : 0;
                            if (MODE == 1) xo &= 0xFFFFFF;
                            if (tan) {
                                const TanEntry<R>& e = *reinterpret_cast<const TanEntry<R>*>(reinterpret_cast<const char*>(cst.data()) + xo);
                                tx[u] = e.t; ax[u] = e.al; bx[u] = e.be;
                                cx[u] = sx[u] = 0;
                            } else {
                                const CS<R>& e = *reinterpret_cast<const CS<R>*>(reinterpret_cast<const char*>(cs.data()) + xo);
                                cx[u] = e.c; sx[u] = e.s;
                                tx[u] = ax[u] = bx[u] = 0;
                            }
                        }
                        A init[4];
                        for (int q = 0; q < 4; ++q) set_amp(init[q], q == 0 ? 1.0 : 0.0);
                        if (tan) {
                            TanRun<A, R, U, 16>::go(D, init, tx, ax, bx, cw, sw, acc[r]);
                        } else {
                            A part = evolve_blocks<A, R, U, MODE>(init, cx, sx, cw, sw, deg, D);
                            if (lay.passes == 1) acc[r] = part; else add_amp(acc[r], part);
                        }
                    }
                }
                for (int m = G_r >> 1; m >= 1; m >>= 1) {
                    std::vector<A> nxt(acc);
                    for (int r = 0; r < G_r; ++r) add_amp(nxt[r], acc[r ^ m]);
                    acc = nxt;
                }
                if (b < K) {
                    out[s * K + b] = (double)acc[0].re / ((double)N * (D + 1));
                    if (amps) {
                        amps[2 * (s * K + b)] = (double)acc[0].re * amp_scale;
                        if constexpr (A::is_complex) amps[2 * (s * K + b) + 1] = (double)acc[0].im * amp_scale;
                        else amps[2 * (s * K + b) + 1] = 0.0;
                    }
                }
            }
        }
    }
    return 0;
}

// ---- a-major kernels (qkan_amajor.cuh): planner, tables, CHEB per input element, SELECT per block, read-out; lane by lane
template <class A, typename R, int DT> struct ChebRun {
    static void go(int D, const A (&init)[4], R c, A& lo0, A& lo2) {
        if (D == DT) cheb_element<A, R, DT>(init, c, lo0, lo2);
        else ChebRun<A, R, DT - 1>::go(D, init, c, lo0, lo2);
    }
};
template <class A, typename R> struct ChebRun<A, R, 0> {
    static void go(int, const A (&)[4], R, A&, A&) {}
};
template <class A, typename R, int DT> struct SelectRun {
    static void go(int D, const A (&lo0)[1], const A (&lo2)[1], const CS<R>* wp, int G, A (&acc)[1]) {
        if (D == DT) select_blocks<A, R, 1, DT>(lo0, lo2, wp, G, acc);
        else SelectRun<A, R, DT - 1>::go(D, lo0, lo2, wp, G, acc);
    }
};
template <class A, typename R> struct SelectRun<A, R, 0> {
    static void go(int, const A (&)[1], const A (&)[1], const CS<R>*, int, A (&)[1]) {}
};

// window != 0: the window kernel's tables and per-row-step input windows
template <class A, typename R>
int emu_amajor(const double* x, const double* W, long long B, int N, int K, int D, int min_g, int max_gk, int window_mode,
               double* out, double* amps) {
    if (D < TAN_MIN_DT || D > TAN_MAX_DT) return -4;
    const BlockLayout lay = plan_amajor_layout(N, K, min_g, max_gk);
    const int G_r = 1 << lay.g_r_log2, G_k = 1 << lay.g_k_log2, G = G_r * G_k;
    if (G < (1 << min_g)) return -3;
    const long long steps = amajor_steps(lay);
    const int window = window_mode ? block_window_max(N, K, lay.g_k_log2, lay.brows) : 0;
    std::vector<CS<R>> wtab((size_t)steps * (D + 1));
    std::vector<int> xotab(steps);
    for (long long e = 0; e < steps; ++e)
        fill_amajor_step<R>(e, W, N, K, D, lay.passes, lay.brows, lay.g_r_log2, lay.g_k_log2, wtab.data(), xotab.data(),
                            (int)sizeof(A), window);
    int NA = 0, NB = 0, L = 0;
    while ((1 << NA) < N) ++NA;
    while ((1 << NB) < K) ++NB;
    while ((1 << L) < D + 1) ++L;
    const double amp_scale = std::pow(2.0, -0.5 * (NA + NB + 2 * L + NA));
    A init[4];
    for (int q = 0; q < 4; ++q) set_amp(init[q], q == 0 ? 1.0 : 0.0);
    // a sample's cs row as the kernels lay it out: plane lo0[0 .. n1), plane lo2[0 .. n1), padded stride
    const int n1 = (window ? window : N) + 1;
    const int row_amps = amajor_row_amps(n1, G, (int)sizeof(A));
    if (row_amps < 2 * n1) return -7;
    const int plane = n1 * (int)sizeof(A);
    for (long long s = 0; s < B; ++s) {
        std::vector<A> full0(N + 1), full2(N + 1);
        for (int n = 0; n <= N; ++n) ChebRun<A, R, 16>::go(D, init, n < N ? clip_unit<R>(x[s * N + n]) : R(0), full0[n], full2[n]);
        for (int bi = 0; bi < lay.brows; ++bi) {
            std::vector<A> row(row_amps);
            for (auto& v : row) set_amp(v, 7.0);             // stale entries must never be read
            if (window) {                                    // the row step's window of the input row, dummy at index `window`
                int lo, len;
                block_window(N, K, lay.g_k_log2, bi, &lo, &len);
                if (len > window) return -6;
                for (int j = 0; j < len; ++j) { row[j] = full0[lo + j]; row[n1 + j] = full2[lo + j]; }
                row[window] = full0[N]; row[n1 + window] = full2[N];
            } else {
                for (int n = 0; n <= N; ++n) { row[n] = full0[n]; row[n1 + n] = full2[n]; }
            }
            const char* rowb = reinterpret_cast<const char*>(row.data());
            for (int k = 0; k < G_k; ++k) {
                const int b = bi * G_k + k;
                std::vector<A> acc(G_r);
                for (int r = 0; r < G_r; ++r) {
                    A a1[1];
                    set_amp(a1[0], 0.0);
                    const int g = k * G_r + r;
                    for (int pi = 0; pi < lay.passes; ++pi) {
                        const size_t st = ((size_t)bi * lay.passes + pi) * G + g;
                        const int xo = xotab[st];
                        if (xo < 0 || xo + (int)sizeof(A) > plane) return -8;
                        A lo0[1], lo2[1];
                        lo0[0] = *reinterpret_cast<const A*>(rowb + xo);
                        lo2[0] = *reinterpret_cast<const A*>(rowb + plane + xo);
                        SelectRun<A, R, 16>::go(D, lo0, lo2, wtab.data() + ((size_t)bi * lay.passes + pi) * (D + 1) * G + g, G, a1);
                    }
                    acc[r] = a1[0];
                }
                for (int m = G_r >> 1; m >= 1; m >>= 1) {
                    std::vector<A> nxt(acc);
                    for (int r = 0; r < G_r; ++r) add_amp(nxt[r], acc[r ^ m]);
                    acc = nxt;
                }
                if (b < K) {
                    out[s * K + b] = (double)acc[0].re / ((double)N * (D + 1));
                    if (amps) {
                        amps[2 * (s * K + b)] = (double)acc[0].re * amp_scale;
                        if constexpr (A::is_complex) amps[2 * (s * K + b) + 1] = (double)acc[0].im * amp_scale;
                        else amps[2 * (s * K + b) + 1] = 0.0;
                    }
                }
            }
        }
    }
    return 0;
}

// direct kernel (every output row reads one input element): lane k of a sample evaluates x[k N / K] and its row's blocks
template <class A, typename R>
int emu_direct(const double* x, const double* W, long long B, int N, int K, int D, double* out, double* amps) {
    if (D < TAN_MIN_DT || D > TAN_MAX_DT) return -4;
    const BlockLayout lay = plan_amajor_layout(N, K, 0);
    if (!amajor_direct_ok(N, K, lay)) return -9;
    const int G = 1 << lay.g_k_log2;
    const long long steps = amajor_steps(lay);
    std::vector<CS<R>> wtab((size_t)steps * (D + 1));
    std::vector<int> xotab(steps);
    for (long long e = 0; e < steps; ++e)
        fill_amajor_step<R>(e, W, N, K, D, lay.passes, lay.brows, lay.g_r_log2, lay.g_k_log2, wtab.data(), xotab.data(), (int)sizeof(A), 0);
    int NA = 0, NB = 0, L = 0;
    while ((1 << NA) < N) ++NA;
    while ((1 << NB) < K) ++NB;
    while ((1 << L) < D + 1) ++L;
    const double amp_scale = std::pow(2.0, -0.5 * (NA + NB + 2 * L + NA));
    A init[4];
    for (int q = 0; q < 4; ++q) set_amp(init[q], q == 0 ? 1.0 : 0.0);
    for (long long s = 0; s < B; ++s)
        for (int k = 0; k < K; ++k) {
            const int nk = (int)(((long long)k * N) / K);
            A lo0[1], lo2[1], acc[1];
            ChebRun<A, R, 16>::go(D, init, clip_unit<R>(x[s * N + nk]), lo0[0], lo2[0]);
            set_amp(acc[0], 0.0);
            for (int a = 0; a < N; ++a) SelectRun<A, R, 16>::go(D, lo0, lo2, wtab.data() + (size_t)a * (D + 1) * G + k, G, acc);
            out[s * K + k] = (double)acc[0].re / ((double)N * (D + 1));
            if (amps) {
                amps[2 * (s * K + k)] = (double)acc[0].re * amp_scale;
                if constexpr (A::is_complex) amps[2 * (s * K + k) + 1] = (double)acc[0].im * amp_scale;
                else amps[2 * (s * K + k) + 1] = 0.0;
            }
        }
    return 0;
}
extern "C" int qkan_emu_direct_forward(int amp, const double* x, const double* W, long long B, int N, int K, int D, double* out, double* amps) {
    if (amp == 0) return emu_direct<Cplx<double>, double>(x, W, B, N, K, D, out, amps);
    if (amp == 1) return emu_direct<Cplx<float>, float>(x, W, B, N, K, D, out, amps);
    if (amp == 2) return emu_direct<Real<double>, double>(x, W, B, N, K, D, out, amps);
    return -2;
}

// element-owner kernel (wide input rows): lanes own input elements; tables, element walk, padding, butterfly
template <class A, typename R>
int emu_elem(const double* x, const double* W, long long B, int N, int K, int D, int min_g, double* out, double* amps, long long* counted) {
    if (D < TAN_MIN_DT || D > TAN_MAX_DT) return -4;
    const ElemLayout lay = plan_elem_layout(N, K, min_g);
    const int G_r = 1 << lay.g_r_log2, G_k = 1 << lay.g_k_log2, G = G_r * G_k;
    const long long steps = elem_steps(lay);
    std::vector<CS<R>> we((size_t)lay.brows * lay.passes * K * (D + 1) * G);
    std::vector<int> xe(steps);
    for (long long e = 0; e < steps; ++e)
        fill_elem_step<R>(e, W, N, K, D, lay.passes, lay.brows, lay.g_r_log2, lay.g_k_log2, we.data(), xe.data());
    int NA = 0, NB = 0, L = 0;
    while ((1 << NA) < N) ++NA;
    while ((1 << NB) < K) ++NB;
    while ((1 << L) < D + 1) ++L;
    const double amp_scale = std::pow(2.0, -0.5 * (NA + NB + 2 * L + NA));
    A init[4];
    for (int q = 0; q < 4; ++q) set_amp(init[q], q == 0 ? 1.0 : 0.0);
    long long cnt = 0;
    for (long long s = 0; s < B; ++s)
        for (int bi = 0; bi < lay.brows; ++bi)
            for (int k = 0; k < G_k; ++k) {
                const int b = bi * G_k + k;
                std::vector<A> acc(G_r);
                for (int r = 0; r < G_r; ++r) {
                    A a1[1];
                    set_amp(a1[0], 0.0);
                    const int g = k * G_r + r;
                    for (int pi = 0; pi < lay.passes; ++pi) {
                        const size_t t = (size_t)bi * lay.passes + pi;
                        const int en = xe[t * G + g];
                        const int n = en < 0 ? 0 : (en & 0xFFFFF);
                        if (n >= N) return -8;
                        if (s == 0 && en >= 0 && (en & (1 << 30))) ++cnt;
                        A lo0[1], lo2[1];
                        ChebRun<A, R, 16>::go(D, init, clip_unit<R>(x[s * N + n]), lo0[0], lo2[0]);
                        for (int j = 0; j < K; ++j) SelectRun<A, R, 16>::go(D, lo0, lo2, we.data() + ((t * K + j) * (D + 1)) * G + g, G, a1);
                    }
                    acc[r] = a1[0];
                }
                for (int m = G_r >> 1; m >= 1; m >>= 1) {
                    std::vector<A> nxt(acc);
                    for (int r = 0; r < G_r; ++r) add_amp(nxt[r], acc[r ^ m]);
                    acc = nxt;
                }
                if (b < K) {
                    out[s * K + b] = (double)acc[0].re / ((double)N * (D + 1));
                    if (amps) {
                        amps[2 * (s * K + b)] = (double)acc[0].re * amp_scale;
                        if constexpr (A::is_complex) amps[2 * (s * K + b) + 1] = (double)acc[0].im * amp_scale;
                        else amps[2 * (s * K + b) + 1] = 0.0;
                    }
                }
            }
    if (counted) *counted = cnt;                              // elements flagged "count the range violation here": must be N
    return 0;
}
extern "C" int qkan_emu_elem_forward(int amp, int min_g, const double* x, const double* W, long long B, int N, int K, int D, double* out,
                                     double* amps, long long* counted) {
    if (amp == 0) return emu_elem<Cplx<double>, double>(x, W, B, N, K, D, min_g, out, amps, counted);
    if (amp == 1) return emu_elem<Cplx<float>, float>(x, W, B, N, K, D, min_g, out, amps, counted);
    if (amp == 2) return emu_elem<Real<double>, double>(x, W, B, N, K, D, min_g, out, amps, counted);
    return -2;
}
extern "C" void qkan_emu_elem_layout(int N, int K, int min_g, int* out4, double* eff) {
    const ElemLayout l = plan_elem_layout(N, K, min_g);
    out4[0] = l.g_r_log2; out4[1] = l.g_k_log2; out4[2] = l.passes; out4[3] = l.brows;
    *eff = l.efficiency;
}

extern "C" int qkan_emu_amajor_forward(int amp, int min_g, int max_gk, int window_mode, const double* x, const double* W,
                                       long long B, int N, int K, int D, double* out, double* amps) {
    if (amp == 0) return emu_amajor<Cplx<double>, double>(x, W, B, N, K, D, min_g, max_gk, window_mode, out, amps);
    if (amp == 1) return emu_amajor<Cplx<float>, float>(x, W, B, N, K, D, min_g, max_gk, window_mode, out, amps);
    if (amp == 2) return emu_amajor<Real<double>, double>(x, W, B, N, K, D, min_g, max_gk, window_mode, out, amps);
    return -2;
}
extern "C" void qkan_emu_amajor_layout(int N, int K, int min_g, int max_gk, int* out4, double* eff) {
    const BlockLayout l = plan_amajor_layout(N, K, min_g, max_gk);
    out4[0] = l.g_r_log2; out4[1] = l.g_k_log2; out4[2] = l.passes; out4[3] = l.brows;
    *eff = l.efficiency;
}
extern "C" int qkan_emu_amajor_row_amps(int n1, int G, int amp_bytes) { return amajor_row_amps(n1, G, amp_bytes); }
extern "C" long long qkan_emu_amajor_smem(int N, int SPC, int row_bytes, int SU, int sub) { return (long long)amajor_smem_bytes(N, SPC, row_bytes, SU, sub); }

// amp: 0 c128, 1 c64, 2 r64
extern "C" int qkan_emu_block_forward(int amp, int mode, int min_g, int tan, const double* x, const double* W, long long B, int N,
                                      int K, int D, double* out, double* amps) {
    const BlockLayout lay = plan_block_layout(N, K, D, min_g);
#define BCASE(AMP, A, R, MODE, UU) \
    if (amp == AMP && mode == MODE && lay.U == UU) return emu_block<A<R>, R, UU, MODE>(x, W, B, N, K, D, min_g, tan, out, amps);
    BCASE(0, Cplx, double, 0, 4) BCASE(0, Cplx, double, 0, 2) BCASE(0, Cplx, double, 0, 1)
    BCASE(0, Cplx, double, 1, 4) BCASE(0, Cplx, double, 1, 2) BCASE(0, Cplx, double, 1, 1)
    BCASE(1, Cplx, float, 0, 4) BCASE(1, Cplx, float, 0, 2) BCASE(1, Cplx, float, 0, 1)
    BCASE(2, Real, double, 0, 4) BCASE(2, Real, double, 0, 2) BCASE(2, Real, double, 0, 1)
    return -2;
}
// host helpers of the block engine, exported for property tests
extern "C" int qkan_emu_tan_row_words(int N, int G, int word_bytes) { return tan_row_words(N, G, word_bytes); }
extern "C" int qkan_emu_cs_row_stride(int N, int G, int pair_bytes) { return cs_row_stride(N, G, pair_bytes); }
extern "C" void qkan_emu_block_window(int N, int K, int g_k_log2, int bi, int* lo, int* len) { block_window(N, K, g_k_log2, bi, lo, len); }
extern "C" int qkan_emu_block_window_max(int N, int K, int g_k_log2, int brows) { return block_window_max(N, K, g_k_log2, brows); }

extern "C" void qkan_emu_block_layout(int N, int K, int D, int min_g, int* out6, double* eff) {
    const BlockLayout l = plan_block_layout(N, K, D, min_g);
    out6[0] = l.U; out6[1] = l.g_r_log2; out6[2] = l.g_k_log2; out6[3] = l.passes; out6[4] = l.brows;
    *eff = l.efficiency;
}

// plan dump for debugging / DESIGN.md
template <int L, int NAT, int NBT, int T, int FW, int PREP = 1> void dump() {
    constexpr auto p = make_plan<L, NAT, NBT, T, FW, PREP>();
    using P = Plan<L, NAT, NBT, T, FW, PREP>;
    printf("plan L=%d NAT=%d NBT=%d T=%d FW=%d PREP=%d: QT=%d stages=%d mstage=%d\n", L, NAT, NBT, T, FW, PREP, P::QT, p.ns, p.mstage);
    for (int s = 0; s < p.ns; ++s) {
        printf("  stage %d local=[", s);
        for (int k = 0; k < T; ++k) printf("%d%s", p.local[s][k], k + 1 < T ? "," : "");
        printf("] h1=%x h3=%x lanes=[", p.h1[s], p.h3[s]);
        for (int k = 0; k < P::NL; ++k) printf("%d%s", p.lane[s][k], k + 1 < P::NL ? "," : "");
        printf("]\n");
    }
}
extern "C" void qkan_emu_dump_plans() {
    dump<2, 2, 2, 4, 3>();
    dump<2, 2, 2, 4, 3, 0>();
    dump<2, 2, 2, 5, 3>();
    dump<5, 3, 3, 5, 3>();
    dump<5, 3, 3, 4, 3>();
    dump<4, 4, 4, 5, 3>();
    dump<4, 4, 2, 4, 3>();
    dump<3, 7, 0, 4, 3>();
    dump<1, 3, 3, 4, 3>();
}
