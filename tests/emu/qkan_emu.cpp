// Host emulation of the CUDA kernel's stage machine (TEST TOOL, not product code).
// Compiles qkan_core.cuh with g++ and runs the thread groups sequentially: inside a
// stage every thread reads and writes only its own amplitudes, and stages are separated
// by barriers on the GPU, so sequential in-place execution is equivalent.  Lets the CPU
// test-suite check the compile-time plans, swizzle and read-out against the oracle.
#include "../../qkan_implementation_b200/csrc/qkan_core.cuh"
#include <vector>
#include <cstring>
#include <cstdio>

using namespace qkan;

template <class A, typename R, class P, int MODE, int S>
struct StageLoop {
    static void run(A* state, const TileArgs<R>& ta, A* acc) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        if constexpr (S < p.ns) {
            for (unsigned t = 0; t < (1u << P::NL); ++t) run_stage<A, R, P, S, MODE>(state, t, ta, acc);
            StageLoop<A, R, P, MODE, S + 1>::run(state, ta, acc);
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
    for (long long s = 0; s < B; ++s) {
        for (auto& a : acc) set_amp(a, 0.0);
        for (int bh = 0; bh < n_bhi; ++bh)
            for (int ah = 0; ah < n_ahi; ++ah) {
                if (sector_is_padding(ah, bh, NAT, NBT, N, K)) continue;
                TileArgs<R> ta{wtab.data(), xidx.data(), x + s * N, NA, ah, bh, D};
                if (PREP == 0) for (size_t i = 0; i < state.size(); ++i) set_amp(state[i], i == 0 ? 1.0 : 0.0);
                StageLoop<A, R, P, MODE, 0>::run(state.data(), ta, acc.data());
            }
        for (int b = 0; b < K; ++b) {
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
