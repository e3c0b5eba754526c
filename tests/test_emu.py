"""CPU: the kernel's compile-time stage plans, swizzle and read-out (qkan_core.cuh compiled with
g++ and run group-by-group) against the oracle.  Config ids are listed in tests/emu/qkan_emu.cpp;
id + 100 = the same tile with the initial Hadamards executed as gates."""
import numpy as np
import pytest

from oracle import qkan_oracle as o

CASES = {0: (4, 4, 3), 1: (4, 4, 3), 2: (4, 4, 3), 3: (4, 4, 3), 4: (4, 4, 3), 5: (4, 4, 3), 6: (4, 4, 3),
         7: (4, 4, 3), 8: (16, 16, 8), 9: (16, 16, 8), 10: (8, 8, 16), 11: (8, 8, 1), 12: (3, 2, 4),
         13: (5, 3, 1), 14: (1, 1, 0), 15: (784, 10, 5), 16: (8, 8, 5), 17: (8, 8, 5), 18: (2, 2, 1),
         19: (4, 4, 10)}
PAPER = (7, 17)
C64 = (5,)


@pytest.mark.parametrize("prep_gates", [0, 1])
@pytest.mark.parametrize("cfg", sorted(CASES))
def test_emulated_kernel_matches_oracle(emu, cfg, prep_gates):
    N, K, D = CASES[cfg]
    rng = np.random.default_rng(cfg)
    B = 2 if N > 100 else 5
    x = rng.uniform(-1, 1, (B, N))
    x[0, 0] = 1.3 if N > 1 else 0.2          # clipped
    if B > 2:
        x[1] = 0.0
        x[2] = 1.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    out = np.zeros((B, K))
    amps = np.zeros((B, K, 2))
    rc = emu.qkan_emu_forward(cfg + 100 * prep_gates, x.ctypes.data, W.ctypes.data, B, N, K, D,
                              out.ctypes.data, amps.ctypes.data)
    assert rc == 0
    ref = o.forward_closed_form(x, W, N, K, D, "paper" if cfg in PAPER else "compat")
    tol = 1e-5 if cfg in C64 else 1e-14
    assert np.abs(out - ref).max() <= tol
    spec = o.circuit_spec(N, K, D)
    assert np.abs(amps[..., 0] * spec.out_scale - ref).max() <= tol
    assert np.abs(amps[..., 1]).max() == 0.0


BLOCK_SHAPES = [(4, 4, 3), (4, 8, 2), (8, 4, 2), (3, 2, 4), (8, 8, 5), (5, 3, 1), (16, 16, 8), (8, 8, 1), (8, 8, 16),
                (4, 4, 10), (1, 1, 0), (2, 2, 1), (1, 5, 2), (7, 1, 3), (33, 3, 2), (100, 10, 5), (4, 4, 40), (784, 10, 5)]


@pytest.mark.parametrize("min_g", [0, 3, 5])
@pytest.mark.parametrize("N,K,D", BLOCK_SHAPES)
def test_emulated_block_engine_matches_oracle(emu, N, K, D, min_g):
    """qkan_block.cuh (layout planner, table entries, block evolution, xor-butterfly read-out) run lane by lane."""
    import ctypes
    f = emu.qkan_emu_block_forward
    f.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
    rng = np.random.default_rng(N * 31 + K * 7 + D + min_g)
    B = 3
    x = rng.uniform(-1.2, 1.2, (B, N))
    x[1] = 0.0
    if N >= 4:
        x[2, :4] = [np.sqrt(0.5), -np.sqrt(0.5), 1e-300, -1.0]      # the quarter-turn boundary, a tiny and a unit input
    W = rng.uniform(-1, 1, (D + 1, N * K))
    spec = o.circuit_spec(N, K, D)
    cases = [(0, 0, 0, 1e-14), (0, 1, 0, 1e-14), (1, 0, 0, 1e-5), (2, 0, 0, 1e-14)]
    for amp, mode, tan, tol in cases:
        out = np.zeros((B, K))
        amps = np.zeros((B, K, 2))
        rc = f(amp, mode, min_g, tan, x.ctypes.data, W.ctypes.data, B, N, K, D, out.ctypes.data, amps.ctypes.data)
        assert rc == 0
        ref = o.forward_closed_form(x, W, N, K, D, "paper" if mode else "compat")
        assert np.abs(out - ref).max() <= tol
        assert np.abs(amps[..., 0] * spec.out_scale - ref).max() <= tol
        assert np.abs(amps[..., 1]).max() == 0.0


@pytest.mark.parametrize("min_g,max_gk", [(0, 5), (3, 5), (5, 5), (0, 0), (2, 1)])
@pytest.mark.parametrize("N,K,D", [s for s in BLOCK_SHAPES if 1 <= s[2] <= 16] + [(6, 6, 7), (9, 2, 12), (64, 5, 16)])
def test_emulated_amajor_kernels_match_oracle(emu, N, K, D, min_g, max_gk):
    """qkan_amajor.cuh (layout planner, a-major tables, one CHEB evolution per (a, b) + SELECT per degree copy,
    xor-butterfly read-out; main and window tables) run lane by lane."""
    import ctypes
    f = emu.qkan_emu_amajor_forward
    f.argtypes = [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
    rng = np.random.default_rng(N * 31 + K * 7 + D + min_g + 3 * max_gk)
    B = 3
    x = rng.uniform(-1.2, 1.2, (B, N))
    x[1] = 0.0
    if N >= 4:
        x[2, :4] = [np.sqrt(0.5), -np.sqrt(0.5), 1e-300, -1.0]      # the quarter-turn boundary, a tiny and a unit input
    W = rng.uniform(-1, 1, (D + 1, N * K))
    spec = o.circuit_spec(N, K, D)
    ref = o.forward_closed_form(x, W, N, K, D, "compat")
    for amp, tol in ((0, 1e-14), (1, 1e-5), (2, 1e-14)):
        for window in (0, 1):
            out = np.zeros((B, K))
            amps = np.zeros((B, K, 2))
            rc = f(amp, min_g, max_gk, window, x.ctypes.data, W.ctypes.data, B, N, K, D, out.ctypes.data, amps.ctypes.data)
            assert rc == 0
            assert np.abs(out - ref).max() <= tol
            assert np.abs(amps[..., 0] * spec.out_scale - ref).max() <= tol
            assert np.abs(amps[..., 1]).max() == 0.0


@pytest.mark.parametrize("N,K,D", [(4, 4, 3), (8, 8, 1), (8, 8, 16), (16, 16, 8), (4, 8, 2), (2, 8, 5), (1, 4, 3), (1, 1, 1), (32, 32, 2), (4, 16, 7)])
def test_emulated_direct_kernel_matches_oracle(emu, N, K, D):
    """Direct kernel (rows that read one input element, K a multiple of N): element index, table walk, read-out."""
    import ctypes
    f = emu.qkan_emu_direct_forward
    f.argtypes = [ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 2
    rng = np.random.default_rng(N + 5 * K + D)
    B = 4
    x = rng.uniform(-1.2, 1.2, (B, N))
    x[1] = 0.0
    W = rng.uniform(-1, 1, (D + 1, N * K))
    spec = o.circuit_spec(N, K, D)
    ref = o.forward_closed_form(x, W, N, K, D, "compat")
    for amp, tol in ((0, 1e-14), (1, 1e-5), (2, 1e-14)):
        out = np.zeros((B, K))
        amps = np.zeros((B, K, 2))
        assert f(amp, x.ctypes.data, W.ctypes.data, B, N, K, D, out.ctypes.data, amps.ctypes.data) == 0
        assert np.abs(out - ref).max() <= tol
        assert np.abs(amps[..., 0] * spec.out_scale - ref).max() <= tol
        assert np.abs(amps[..., 1]).max() == 0.0
    assert f(0, x.ctypes.data, W.ctypes.data, B, 3, 4, D, out.ctypes.data, amps.ctypes.data) == -9      # K not a multiple of N


@pytest.mark.parametrize("min_g", [0, 2, 5])
@pytest.mark.parametrize("N,K,D", [(784, 10, 5), (100, 10, 5), (33, 3, 2), (7, 1, 3), (8, 4, 2), (5, 8, 1), (4, 4, 3), (600, 16, 4), (13, 7, 16), (1, 5, 2), (64, 64, 1)])
def test_emulated_element_owner_kernel_matches_oracle(emu, N, K, D, min_g):
    """Element-owner walk (qkan_block_elem_kernel): layout planner, element / SELECT tables with padding, one range count per
    input, xor-butterfly read-out; lane by lane."""
    import ctypes
    f = emu.qkan_emu_elem_forward
    f.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_longlong] + [ctypes.c_int] * 3 + [ctypes.c_void_p] * 3
    rng = np.random.default_rng(N * 13 + K * 5 + D + min_g)
    B = 2
    x = rng.uniform(-1.2, 1.2, (B, N))
    W = rng.uniform(-1, 1, (D + 1, N * K))
    spec = o.circuit_spec(N, K, D)
    ref = o.forward_closed_form(x, W, N, K, D, "compat")
    for amp, tol in ((0, 1e-14), (1, 1e-5), (2, 1e-14)):
        out = np.zeros((B, K))
        amps = np.zeros((B, K, 2))
        counted = ctypes.c_longlong(-1)
        assert f(amp, min_g, x.ctypes.data, W.ctypes.data, B, N, K, D, out.ctypes.data, amps.ctypes.data, ctypes.byref(counted)) == 0
        assert np.abs(out - ref).max() <= tol
        assert np.abs(amps[..., 0] * spec.out_scale - ref).max() <= tol
        assert np.abs(amps[..., 1]).max() == 0.0
        assert counted.value == N                      # every input's range violation is reported by exactly one (row, lane)


def test_amajor_row_strides_spread_the_lanes(emu):
    """amajor_row_amps: the lanes of one shared-memory phase (128 bytes) - P / G sample rows x G consecutive amplitudes of a
    plane - land on different slots for the small power-of-two groups; rows hold both planes."""
    for n1, G in ((5, 4), (9, 8), (17, 16), (6, 4), (4, 2), (13, 8), (3, 1), (800, 16)):
        for ab in (16, 8):
            P = 128 // ab
            rs = emu.qkan_emu_amajor_row_amps(n1, G, ab)
            assert rs >= 2 * n1
            if G < P and n1 <= 2 * P and G <= n1:
                slots = [((lane // G) * rs + lane % G) % P for lane in range(P)]
                assert len(set(slots)) == P, (n1, G, ab, rs)


def test_row_strides_are_bank_conflict_free(emu):
    """tan_row_words / cs_row_stride: the lanes of one shared-memory phase (128 bytes) that read the same member of G
    consecutive entries in P / G consecutive sample rows must land on different banks (G = 4, 8, 16 in FP64)."""
    for N, G in ((4, 4), (8, 8), (16, 16), (5, 4), (3, 2), (12, 8)):
        for wb in (8, 4):
            P = 128 // wb
            rsw = emu.qkan_emu_tan_row_words(N, G, wb)
            assert rsw >= 3 * (N + 1)
            banks = [((lane // G) * rsw + 3 * (lane % G)) % P for lane in range(P)] if G < P else list(range(P))
            worst = max(banks.count(b) for b in set(banks))
            assert worst == 1 or G not in (4, 8, 16), (N, G, wb, rsw, worst)
        np16 = emu.qkan_emu_cs_row_stride(N, G, 16)
        assert np16 >= N + 1
        if G < 8:
            banks = [((lane // G) * np16 + lane % G) % 8 for lane in range(8)]
            assert len(set(banks)) == 8


def test_input_windows_cover_every_block(emu):
    """block_window: row step bi reads x[(a + N b) / K] for its rows b only inside [lo, lo + len), windows are
    monotone, and block_window_max bounds them."""
    import ctypes
    for N, K, gk in ((784, 10, 0), (784, 10, 1), (100, 10, 0), (33, 3, 1), (600, 16, 2), (7, 5, 0), (5, 8, 1), (1000, 3, 0)):
        Gk = 1 << gk
        brows = (K + Gk - 1) // Gk
        wmax = emu.qkan_emu_block_window_max(N, K, gk, brows)
        prev_lo = -1
        for bi in range(brows):
            lo, ln = ctypes.c_int(), ctypes.c_int()
            emu.qkan_emu_block_window(N, K, gk, bi, ctypes.byref(lo), ctypes.byref(ln))
            assert 1 <= ln.value <= wmax and lo.value >= prev_lo
            prev_lo = lo.value
            rows = [b for b in range(bi * Gk, min(K, (bi + 1) * Gk))]
            idx = [(a + N * b) // K for b in rows for a in range(N)]
            assert min(idx) == lo.value and max(idx) == lo.value + ln.value - 1
