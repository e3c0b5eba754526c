// qkan_core.cuh - register-tile statevector engine for the QKAN forward circuit.
//
// The circuit acts on the register
//     deg[L] | f_x | f_w | a[NA] | b[NB]          (little-endian amplitude index)
// with gates   H1 = H on every a, b, deg qubit          (SUM pre-layer, PREPARE)
//              CHEB = D x UCRy(theta_x) on f_x, controlled by (a, b)
//              MUL  = UCRy(theta_w) on f_w, controlled by (a, b, deg)
//              H3 = H on every deg and a qubit           (UNPREPARE, SUM)
// A *tile* is the sub-register deg[L] | f_x | f_w | a[NAT] | b[NBT] that one thread
// group keeps on chip; the remaining high a / b qubits (if any) are enumerated as
// sectors by the kernel.  Inside a tile every thread owns 2^T amplitudes in
// registers ("local" qubits); a *stage* = load 2^T amplitudes from shared memory,
// apply every gate whose target is local, store.  The stage list is computed at
// compile time by make_plan().  This header is shared by the CUDA kernel and by
// the host emulation used in the CPU tests (tests/emu), so the index logic is
// checked without a GPU.
#pragma once
#include <stdint.h>
#include <math.h>

#if defined(__CUDACC__)
#define QK_HD __host__ __device__ __forceinline__
#define QK_UNROLL _Pragma("unroll")
#else
#define QK_HD inline
#define QK_UNROLL
#endif

namespace qkan {

// ---------------------------------------------------------------- amplitudes
template <typename R> struct alignas(2 * sizeof(R)) Cplx {
    R re, im;
    static constexpr bool is_complex = true;
};
template <typename R> struct alignas(sizeof(R)) Real {
    R re;
    static constexpr bool is_complex = false;
};

QK_HD double qk_fma(double a, double b, double c) { return fma(a, b, c); }
QK_HD float qk_fma(float a, float b, float c) { return fmaf(a, b, c); }
QK_HD double qk_sqrt(double a) { return sqrt(a); }
QK_HD float qk_sqrt(float a) { return sqrtf(a); }

// un-normalised Hadamard butterfly: (u, v) <- (u + v, u - v); the 1/sqrt(2) of every
// H pass is folded into the single read-out scale.
template <typename R> QK_HD void bfly(Cplx<R>& u, Cplx<R>& v) {
    R a = u.re, b = u.im;
    u.re = a + v.re; u.im = b + v.im;
    v.re = a - v.re; v.im = b - v.im;
}
template <typename R> QK_HD void bfly(Real<R>& u, Real<R>& v) {
    R a = u.re;
    u.re = a + v.re;
    v.re = a - v.re;
}
// Ry pass on one amplitude pair: (u, v) <- (c u - s v, s u + c v); 1 MUL + 1 FMA per real output
template <typename R> QK_HD void rot(Cplx<R>& u, Cplx<R>& v, R c, R s) {
    R ur = u.re, ui = u.im, vr = v.re, vi = v.im;
    u.re = qk_fma(c, ur, -(s * vr));
    u.im = qk_fma(c, ui, -(s * vi));
    v.re = qk_fma(s, ur, c * vr);
    v.im = qk_fma(s, ui, c * vi);
}
template <typename R> QK_HD void rot(Real<R>& u, Real<R>& v, R c, R s) {
    R ur = u.re, vr = v.re;
    u.re = qk_fma(c, ur, -(s * vr));
    v.re = qk_fma(s, ur, c * vr);
}
template <typename A> QK_HD void set_amp(A& a, double re) {
    a.re = (decltype(a.re))re;
    if constexpr (A::is_complex) a.im = 0;
}
template <typename A> QK_HD void add_amp(A& a, const A& b) {
    a.re += b.re;
    if constexpr (A::is_complex) a.im += b.im;
}

// ---------------------------------------------------------------------- plan
// PREP = 1: the state preparation H^(x)(m+l) |0...0> (uniform product state) is written in
//           closed form and the stage list starts with the multiplexor stage;
// PREP = 0: the state starts as |0...0> in shared memory and the initial Hadamards are
//           executed as butterfly passes like every other gate ("gates" preparation).
template <int L_, int NAT_, int NBT_, int T_, int FW_, int PREP_>
struct Plan {
    static constexpr int L = L_, NAT = NAT_, NBT = NBT_, T = T_, FW = FW_, PREP = PREP_;
    static constexpr int QT = L + 2 + NAT + NBT;      // tile qubits
    static constexpr int NL = QT - T;                 // qubits spread over the thread group
    static constexpr int MAXS = 12;
    static constexpr int BIT_FX = L, BIT_FW = L + 1, BIT_A0 = L + 2, BIT_B0 = L + 2 + NAT;
    int ns = 0;                 // number of stages
    int mstage = 0;             // the stage holding f_x, f_w (all multiplexor passes)
    int local[MAXS][T] = {};    // local qubit positions per stage, ascending
    int lane[MAXS][NL > 0 ? NL : 1] = {};   // thread-index bit i -> amplitude-index bit
    unsigned h1[MAXS] = {};     // bit k set: initial H on local slot k in this stage
    unsigned h3[MAXS] = {};     // bit k set: final H on local slot k in this stage
};

constexpr bool is_deg(int p, int L) { return p < L; }
constexpr bool is_a(int p, int L, int NAT) { return p >= L + 2 && p < L + 2 + NAT; }
constexpr bool is_b(int p, int L, int NAT) { return p >= L + 2 + NAT; }

template <int L, int NAT, int NBT, int T, int FW, int PREP>
constexpr Plan<L, NAT, NBT, T, FW, PREP> make_plan() {
    using P = Plan<L, NAT, NBT, T, FW, PREP>;
    constexpr int QT = P::QT;
    static_assert(T >= 2 && T <= QT, "need 2 <= T <= tile qubits");
    P p{};
    bool inM[QT] = {};
    int mloc[T] = {};
    int nm = 0;
    mloc[nm++] = L;     inM[L] = true;
    mloc[nm++] = L + 1; inM[L + 1] = true;
    for (int i = 0; i < L && nm < T; ++i) { mloc[nm++] = i; inM[i] = true; }
    for (int i = 0; i < NAT && nm < T; ++i) { mloc[nm++] = L + 2 + i; inM[L + 2 + i] = true; }
    for (int i = 0; i < NBT && nm < T; ++i) { mloc[nm++] = L + 2 + NAT + i; inM[L + 2 + NAT + i] = true; }

    int s = 0;
    // ---- phase A: initial H on every a / b / deg qubit that is not local in the M stage
    int pend[QT > 0 ? QT : 1] = {};
    int np_ = 0;
    if (PREP == 0) {
        for (int q = L + 2; q < QT; ++q) if (!inM[q]) pend[np_++] = q;
        for (int q = 0; q < L; ++q) if (!inM[q]) pend[np_++] = q;
    }
    for (int c = 0; c < np_; c += T) {
        int cnt = (np_ - c < T) ? (np_ - c) : T;
        bool used[QT] = {};
        bool want[QT] = {};
        for (int i = 0; i < cnt; ++i) { used[pend[c + i]] = true; want[pend[c + i]] = true; }
        int n = cnt;
        for (int q = QT - 1; q >= 0 && n < T; --q) if (!used[q]) { used[q] = true; ++n; }
        int k = 0;
        for (int q = 0; q < QT; ++q) if (used[q]) { p.local[s][k] = q; if (want[q]) p.h1[s] |= 1u << k; ++k; }
        ++s;
    }
    // ---- phase M: H1 on local index qubits, all multiplexor passes, H3 on local deg / a qubits
    {
        int k = 0;
        for (int q = 0; q < QT; ++q) if (inM[q]) {
            p.local[s][k] = q;
            if (PREP == 0 && q != L && q != L + 1) p.h1[s] |= 1u << k;
            if (is_deg(q, L) || is_a(q, L, NAT)) p.h3[s] |= 1u << k;
            ++k;
        }
        p.mstage = s;
        ++s;
    }
    // ---- phase F: final H on deg / a qubits that were not local in the M stage
    np_ = 0;
    for (int q = 0; q < L; ++q) if (!inM[q]) pend[np_++] = q;
    for (int q = L + 2; q < L + 2 + NAT; ++q) if (!inM[q]) pend[np_++] = q;
    for (int c = 0; c < np_; c += T) {
        int cnt = (np_ - c < T) ? (np_ - c) : T;
        bool used[QT] = {};
        bool want[QT] = {};
        for (int i = 0; i < cnt; ++i) { used[pend[c + i]] = true; want[pend[c + i]] = true; }
        int n = cnt;
        for (int q = QT - 1; q >= 0 && n < T; --q) if (!used[q]) { used[q] = true; ++n; }
        int k = 0;
        for (int q = 0; q < QT; ++q) if (used[q]) { p.local[s][k] = q; if (want[q]) p.h3[s] |= 1u << k; ++k; }
        ++s;
    }
    p.ns = s;
    // ---- thread-bit -> index-bit order per stage.  The first FW thread bits address the
    // 2^FW lanes that share one shared-memory phase (8 x 16 B or 16 x 8 B = 128 B); the
    // swizzle (phys_index) sends index bit q to bank-group bit q % FW, so they must have
    // pairwise different residues to be conflict free.
    for (int st = 0; st < s; ++st) {
        bool loc[QT] = {};
        for (int k = 0; k < T; ++k) loc[p.local[st][k]] = true;
        bool taken[QT] = {};
        bool res[FW] = {};
        int n = 0;
        for (int q = 0; q < QT && n < FW; ++q)
            if (!loc[q] && !res[q % FW]) { res[q % FW] = true; taken[q] = true; if (P::NL > 0) p.lane[st][n] = q; ++n; }
        for (int q = 0; q < QT; ++q)
            if (!loc[q] && !taken[q]) { if (P::NL > 0) p.lane[st][n] = q; ++n; }
    }
    return p;
}

// shared-memory swizzle: XOR every FW-bit group of the index into the lowest group
template <int FW> constexpr QK_HD unsigned fold_bits(unsigned idx) {
    unsigned f = 0;
    for (int k = 1; k * FW < 24; ++k) f ^= (idx >> (k * FW)) & ((1u << FW) - 1u);
    return f;
}
template <int FW> constexpr QK_HD unsigned phys_index(unsigned idx) { return idx ^ fold_bits<FW>(idx); }

// deposit the T bits of j into the local positions of stage S
template <class P> constexpr unsigned dep_local(const P& p, int S, unsigned j) {
    unsigned r = 0;
    for (int k = 0; k < P::T; ++k) r |= ((j >> k) & 1u) << p.local[S][k];
    return r;
}

// ---------------------------------------------------------- per-launch tables
// wtab[(ab << L) | deg] = (cos, sin)(theta_w / 2) = (w, sqrt(1 - w^2)), (0, 1) on padding
// xidx[ab]              = index into the sample's x row, or -1 on padding
// with ab = (b << NA) | a over the padded K x N grid (reference index i = a + N b,
// QKANLayer.py:132; x index i // K, ChebyshevStep.py:64).
template <typename R> struct alignas(2 * sizeof(R)) CS { R c, s; };

template <typename R>
struct TileArgs {
    const CS<R>* wtab;
    const int* xidx;
    const double* xrow;     // this sample's x[N] (shared memory in the kernel)
    int NA;                 // full number of a qubits
    int a_hi, b_hi;         // sector = values of the a / b qubits above the tile
    int D;                  // number of CHEB applications (MODE 0 = compat: all terms; 1 = paper: term d gets d)
    // direct read-out (tile = whole register): rows of out / amps of this sample, or null
    int K;
    double* out_row;
    void* amp_row;
    double out_scale, amp_scale;
};

template <typename R> QK_HD R clip_unit(double x) {
    // one comparison on |x| (not fmin/fmax), so that NaN propagates like np.clip (ChebyshevStep.py:52)
    const double y = fabs(x) > 1.0 ? copysign(1.0, x) : x;
    return (R)y;
}

// ------------------------------------------------------------- stage executor
// state: the tile's 2^QT amplitudes in (swizzled) shared memory;  t: thread in group.
// acc: K_pad accumulators of the post-selected amplitudes (last stage only).
template <class P> constexpr int find_slot(const P& p, int S, int bit) {
    for (int k = 0; k < P::T; ++k) if (p.local[S][k] == bit) return k;
    return -1;
}

// Sample-independent part of the multiplexor stage, per thread: the weight rotation
// (cos, sin) of every local control value and the x-row index feeding the input rotation.
// Depends on the sector only, so the kernel loads it once per sector (once per launch when
// the whole register fits the tile).
template <typename R, class P>
struct MuxCoef {
    static constexpr int NC = P::T - 2;
    R cw[1 << NC], sw[1 << NC];
    int xi[1 << NC];
    int dloc[1 << NC];

    // local index j (flags ignored) -> compressed control index c
    static constexpr int compress(int j) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        constexpr int kx = find_slot(p, p.mstage, P::BIT_FX), kw = find_slot(p, p.mstage, P::BIT_FW);
        int c = 0, cc = 0;
        for (int k = 0; k < P::T; ++k) if (k != kx && k != kw) { c |= ((j >> k) & 1) << cc; ++cc; }
        return c;
    }
    static constexpr unsigned expand(int c) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        constexpr int kx = find_slot(p, p.mstage, P::BIT_FX), kw = find_slot(p, p.mstage, P::BIT_FW);
        unsigned j = 0;
        int cc = 0;
        for (int k = 0; k < P::T; ++k) if (k != kx && k != kw) { j |= ((unsigned)(c >> cc) & 1u) << k; ++cc; }
        return j;
    }
    // the x coefficient depends on (a, b) only: entry that shares it and has all deg slots = 0
    static constexpr int ab_source(int c) {
        constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
        constexpr int kx = find_slot(p, p.mstage, P::BIT_FX), kw = find_slot(p, p.mstage, P::BIT_FW);
        int cc = 0, cab = 0;
        for (int k = 0; k < P::T; ++k) if (k != kx && k != kw) {
            if (!is_deg(p.local[p.mstage][k], P::L)) cab |= ((c >> cc) & 1) << cc;
            ++cc;
        }
        return cab;
    }
};

template <typename R, class P>
QK_HD void load_mux_coefs(MuxCoef<R, P>& mc, unsigned t, const TileArgs<R>& ta) {
    constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
    constexpr int S = p.mstage, L = P::L, NAT = P::NAT, NBT = P::NBT;
    unsigned base = 0;
    QK_UNROLL
    for (int i = 0; i < P::NL; ++i) base |= ((t >> i) & 1u) << p.lane[S][i];
    const unsigned a_thr = (base >> P::BIT_A0) & ((1u << NAT) - 1u);
    const unsigned b_thr = (base >> P::BIT_B0) & ((1u << NBT) - 1u);
    const unsigned d_thr = base & ((1u << L) - 1u);
    const unsigned ab_thr = ((((unsigned)ta.b_hi << NBT) | b_thr) << ta.NA) | (((unsigned)ta.a_hi << NAT) | a_thr);
    QK_UNROLL
    for (int c = 0; c < (1 << MuxCoef<R, P>::NC); ++c) {
        const unsigned dj = dep_local(p, S, MuxCoef<R, P>::expand(c));
        const unsigned a_loc = (dj >> P::BIT_A0) & ((1u << NAT) - 1u);
        const unsigned b_loc = (dj >> P::BIT_B0) & ((1u << NBT) - 1u);
        const unsigned d_loc = dj & ((1u << L) - 1u);
        const unsigned ab = ab_thr + (b_loc << ta.NA) + a_loc;
        const CS<R> w = ta.wtab[(ab << L) | d_thr | d_loc];
        mc.cw[c] = w.c;
        mc.sw[c] = w.s;
        mc.dloc[c] = (int)(d_thr | d_loc);
        mc.xi[c] = ta.xidx[ab];
    }
}

// DIRECT: the tile is the whole register, so the post-selected amplitudes of the last stage go
// straight from registers to global memory (otherwise they are summed over sectors in `acc`).
template <class A, typename R, class P, int S, int MODE, bool DIRECT = false>
QK_HD void run_stage(A* state, unsigned t, const TileArgs<R>& ta, const MuxCoef<R, P>& mc, A* acc) {
    constexpr P p = make_plan<P::L, P::NAT, P::NBT, P::T, P::FW, P::PREP>();
    constexpr int T = P::T, NLOC = 1 << T, L = P::L, NAT = P::NAT, NBT = P::NBT, FW = P::FW;
    constexpr bool last = (S == p.ns - 1), mux = (S == p.mstage);
    // closed-form preparation: the first stage is the multiplexor stage and starts from the
    // (un-normalised) uniform state: 1 wherever f_x = f_w = 0
    constexpr bool first = (P::PREP == 1 && S == 0);

    unsigned base = 0;
    QK_UNROLL
    for (int i = 0; i < P::NL; ++i) base |= ((t >> i) & 1u) << p.lane[S][i];
    const unsigned pt = phys_index<FW>(base);

    A v[NLOC];
    if constexpr (first) {
        QK_UNROLL
        for (int j = 0; j < NLOC; ++j)
            set_amp(v[j], (dep_local(p, S, j) & ((1u << P::BIT_FX) | (1u << P::BIT_FW))) == 0 ? 1.0 : 0.0);
    } else {
        QK_UNROLL
        for (int j = 0; j < NLOC; ++j) v[j] = state[pt ^ phys_index<FW>(dep_local(p, S, j))];
    }

    // ---- initial Hadamards on local slots
    QK_UNROLL
    for (int k = 0; k < T; ++k) {
        if ((p.h1[S] >> k) & 1u) {
            QK_UNROLL
            for (int j = 0; j < NLOC; ++j)
                if (!((j >> k) & 1)) bfly(v[j], v[j | (1 << k)]);
        }
    }

    // ---- multiplexor passes
    if constexpr (mux) {
        constexpr int kx = find_slot(p, S, P::BIT_FX);
        constexpr int kw = find_slot(p, S, P::BIT_FW);
        static_assert(kx >= 0 && kw >= 0, "M stage must hold both flag qubits");
        constexpr int NC = T - 2;                 // local control slots (all non-flag slots)
        // x coefficients of this sample: cos(theta_x/2) = clip(x), sin = sqrt(1 - x^2) (no acos needed)
        R cx[1 << NC], sx[1 << NC];
        QK_UNROLL
        for (int c = 0; c < (1 << NC); ++c) {
            const int src = MuxCoef<R, P>::ab_source(c);
            if (src == c) {
                const int xi = mc.xi[c];
                const R xc = xi >= 0 ? clip_unit<R>(ta.xrow[xi]) : (R)0;
                cx[c] = xc;
                sx[c] = qk_sqrt((R(1) - xc) * (R(1) + xc));
            } else {
                cx[c] = cx[src]; sx[c] = sx[src];
            }
        }
        // CHEB: D applications.  Z . Ry(-theta) . Z (odd applications) equals Ry(+theta) as a
        // matrix, so every application is the same real 2x2 pass.
        for (int r = 0; r < ta.D; ++r) {
            QK_UNROLL
            for (int j = 0; j < NLOC; ++j) {
                if (!((j >> kx) & 1)) {
                    const int c = MuxCoef<R, P>::compress(j);
                    if (MODE == 0 || mc.dloc[c] >= r + 1) rot(v[j], v[j | (1 << kx)], cx[c], sx[c]);
                }
            }
        }
        // MUL / SELECT
        QK_UNROLL
        for (int j = 0; j < NLOC; ++j) {
            if (!((j >> kw) & 1)) {
                const int c = MuxCoef<R, P>::compress(j);
                rot(v[j], v[j | (1 << kw)], mc.cw[c], mc.sw[c]);
            }
        }
    }

    // ---- final Hadamards on local slots
    QK_UNROLL
    for (int k = 0; k < T; ++k) {
        if ((p.h3[S] >> k) & 1u) {
            QK_UNROLL
            for (int j = 0; j < NLOC; ++j)
                if (!((j >> k) & 1)) bfly(v[j], v[j | (1 << k)]);
        }
    }

    if constexpr (last) {
        // post-selection deg = f_x = f_w = a = 0: everything below the b qubits must be 0
        constexpr unsigned nonb = (1u << P::BIT_B0) - 1u;
        if ((base & nonb) == 0) {
            QK_UNROLL
            for (int j = 0; j < NLOC; ++j) {
                const unsigned dj = dep_local(p, S, j);
                if ((dj & nonb) == 0) {
                    const unsigned b = (((unsigned)ta.b_hi << NBT) | ((base | dj) >> P::BIT_B0));
                    if constexpr (DIRECT) {
                        if (ta.out_row != nullptr && (int)b < ta.K) {
                            ta.out_row[b] = (double)v[j].re * ta.out_scale;
                            if (ta.amp_row != nullptr) {
                                Cplx<R> z;
                                z.re = (R)((double)v[j].re * ta.amp_scale);
                                if constexpr (A::is_complex) z.im = (R)((double)v[j].im * ta.amp_scale);
                                else z.im = (R)0;
                                reinterpret_cast<Cplx<R>*>(ta.amp_row)[b] = z;
                            }
                        }
                    } else {
                        add_amp(acc[b], v[j]);
                    }
                }
            }
        }
    } else {
        QK_UNROLL
        for (int j = 0; j < NLOC; ++j) state[pt ^ phys_index<FW>(dep_local(p, S, j))] = v[j];
    }
}

// ---------------------------------------------------------------- table prep
// One entry ab = (b << NA) | a of the padded K x N grid: writes xidx[ab] and the 2^L
// weight entries wtab[(ab << L) | deg].  W is the reference's [D+1, N*K] weight matrix
// (MulStep.py:22; flat index i = a + N b as read by the SUM reshape, QKANLayer.py:132).
template <typename R>
QK_HD void fill_tables_entry(unsigned ab, const double* W, int N, int K, int D, int NA, int L,
                             CS<R>* wtab, int* xidx) {
    const int a = (int)(ab & ((1u << NA) - 1u)), b = (int)(ab >> NA);
    const bool valid = a < N && b < K;
    const int i = a + N * b;
    xidx[ab] = valid ? i / K : -1;
    for (int d = 0; d < (1 << L); ++d) {
        const R w = (valid && d <= D) ? (R)W[(long long)d * N * K + i] : (R)0;
        CS<R> e;
        e.c = w;
        e.s = qk_sqrt((R(1) - w) * (R(1) + w));
        wtab[((size_t)ab << L) | (unsigned)d] = e;
    }
}

// a sector (values of the a / b qubits above the tile) whose rows or columns are all
// padding carries only zero amplitudes into the read-out and is not simulated
QK_HD bool sector_is_padding(int a_hi, int b_hi, int NAT, int NBT, int N, int K) {
    return ((a_hi << NAT) >= N) || ((b_hi << NBT) >= K);
}

// executed-work accounting for one tile (used by bench / DESIGN): passes over 2^QT amplitudes
template <class P> constexpr int tile_h_passes() {
    return P::PREP ? (P::L + P::NAT) : (2 * P::L + 2 * P::NAT + P::NBT);
}

}  // namespace qkan
