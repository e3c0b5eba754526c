"""Writes instances/inst_*.cu and qkan_instances.h: the explicit template instantiations of
qkan_forward_kernel (one translation unit per group so `make -j` builds them in parallel).

    python gen_instances.py

Tile naming: (L, NAT, NBT) = deg / a / b qubits kept on chip; T = qubits per thread.
"""
import os

HERE = os.path.dirname(os.path.abspath(__file__))
AMPS = {"c128": ("Cplx<double>", "double", 16), "c64": ("Cplx<float>", "float", 8), "r64": ("Real<double>", "double", 8)}
MAX_SMEM = 200 * 1024     # leave room for x tiles / accumulators under the 227 KB limit


def shape(amp, L, NAT, NBT, T=None, NT=None, MINB=1):
    """pick T / NT for a tile; returns None if it cannot fit one CTA"""
    QT = L + 2 + NAT + NBT
    if T is None and NT is None:
        # measured on B200 (profiles/r01_tune_variants_2.jsonl): small tiles want few registers and many
        # CTAs, 12-13 qubit tiles want 32 amplitudes per thread
        if QT <= 9:
            T, NT, MINB = min(3, QT), 256, 3
        elif QT <= 11:
            T, NT, MINB = 4, 128, 4
        elif QT == 12:
            T, NT, MINB = 5, 128, 2
        else:
            T, NT, MINB = 5, 256, 1
    if T is None:
        T = min(4, QT)
    T = max(2, min(T, QT))
    G = 1 << (QT - T)
    if NT is None:
        NT = max(256, G)
    if G > NT:
        NT = G
    if G > 1024 or NT > 1024 or G > NT:
        return None
    if NT * MINB > 2048 or 65536 // (NT * MINB) < 40:
        return None
    spi = NT // G
    if G > 32 and G != NT and spi > 15:
        return None
    if spi * (1 << QT) * AMPS[amp][2] > MAX_SMEM:
        return None
    return (T, NT, MINB)


def instances():
    """staged full-statevector kernels (prep = gates): every gate incl. the initial Hadamards is a pass"""
    out = []   # (group, amp, L, NAT, NBT, T, NT, MINB, MODE, PREP, prio, variant, full)
    # generic tiles: b fully in sectors, a in the tile up to 7 qubits -> any shape with D <= 31
    for amp in ("c128",):      # the validation engine: generic tiles in the default precision only
        for L in range(0, 6):
            for NAT in range(0, 8):
                s = shape(amp, L, NAT, 0)
                if s is None:
                    continue
                out.append((f"gen_{amp}_L{L}", amp, L, NAT, 0) + s + (0, 0, 0, 0, "false"))
    # tiles holding the whole register of the BASELINE.json / reference-test shapes
    hot = [
        (2, 2, 2, 1),                                                    # N4 K4 D3      8 qubits
        (1, 3, 3, 1), (2, 3, 3, 1), (3, 3, 3, 1), (4, 3, 3, 1), (5, 3, 3, 1),   # N8 K8 D1..16  9..13 qubits
        (4, 4, 3, 0), (4, 4, 2, 0),                                      # N16 K16 D8    14 qubits -> 13 / 12 qubit tile
        (3, 8, 0, 0),                                                    # N784 K10 D5   19 qubits -> 13 qubit tile
        (2, 2, 3, 1), (2, 3, 2, 1), (3, 2, 2, 1),                        # the reference tests' 4x8, 8x4 shapes, 4x4 d5
    ]
    for amp in ("c128", "c64", "r64"):
        for (L, NAT, NBT, full) in hot:
            s = shape(amp, L, NAT, NBT)
            if s is None:
                continue
            out.append((f"hot_{amp}", amp, L, NAT, NBT) + s + (0, 0, 10, 0, "true" if full else "false"))
    return out


# block engine, generic (cos, sin) kernels (paper mode, D = 0, D > 16, very wide rows): (U, NT) -> min CTAs per SM.
# U = blocks per lane in registers (1 by default, 4 when shared memory limits the resident warps).
BLOCK_MINB = {
    (4, 256): 2, (4, 128): 4, (4, 64): 8, (4, 32): 16,
    (1, 256): 4, (1, 128): 8, (1, 64): 16, (1, 32): 32,
}
DT_MAX = 16     # compile-time degree specialisations (compat mode): D = 1 .. DT_MAX


def block_instances():
    out = []   # (group, amp, U, SU, MODE, NT, MINB, DT, is_default)
    for amp in ("c128", "c64", "r64"):
        for mode in (0, 1):
            for (U, NT), minb in BLOCK_MINB.items():
                out.append((f"block_{amp}_m{mode}", amp, U, 1, mode, NT, minb, 0, 1))
    return out


# a-major scaled-rotation kernels (qkan_amajor.cuh; the default for compat mode, 1 <= D <= DT_MAX):
# (SU, NT) -> [(MINB, is_default)].  SU = samples per lane at a time.  Non-default entries are tuning variants
# (QKAN_BLOCK_TUNE=1:NT:MINB:SU), built for complex128 only.
AMAJOR = {
    (2, 256): [(3, 1)],
    (1, 256): [(4, 1)],      # fall-backs when two samples' rows do not fit (wide input rows)
    (1, 128): [(8, 1)],
}
# direct kernels (every output row reads one input element: K a multiple of N, K a power of two): (SU, NT) -> [(MINB, is_default)]
# measured best on every BASELINE shape: four samples per lane, 256-thread CTAs, two CTAs per SM
# (profiles/r02e_tune_direct.jsonl, r02d_tune_direct_all.jsonl)
DIRECT = {
    (4, 256): [(2, 1)],      # (2, 256) x 3 CTAs, (4, 128) x 4 and (2, 128) x 6 measured equal or slower, with run-time and with
                             # compile-time lanes per sample (profiles/r02e_tune_direct.jsonl, r02u_tune_direct.jsonl), and so did three samples per lane
                             # at 94 registers with 4 or 5 CTAs of 128 threads (20 resident warps: occupancy is not the limiter): not built
}
# complex64 (half the registers per sample): (4, 256) x 3 or 4 CTAs and (8, 256) x 2 or 3 CTAs measured within noise of / slower than
# the default on every BASELINE shape (profiles/r02y_tune_c64.jsonl): not built
DIRECT_EXTRA = {}
# element-owner kernels (wide input rows): (SU, NT) -> [(MINB, is_default)]; no shared memory, so four samples per lane always launch
ELEM = {
    (4, 256): [(2, 1)],
}
WINDOW_CTAS = ()   # the window kernel of round 1 is superseded by the element-owner kernel


def amajor_instances():
    out = []   # (group, amp, SU, NT, MINB, DT, is_default)
    for amp in ("c128", "c64", "r64"):
        for dt in range(1, DT_MAX + 1):
            for (SU, NT), variants in AMAJOR.items():
                for (minb, dflt) in variants:
                    if not dflt and amp != "c128":
                        continue
                    out.append((f"amajor_{amp}_d{dt}", amp, SU, NT, minb, dt, dflt))
    return out


def elem_instances():
    out = []   # (group, amp, SU, NT, MINB, DT, is_default)
    for amp in ("c128", "c64", "r64"):
        for dt in range(1, DT_MAX + 1):
            for (SU, NT), variants in ELEM.items():
                for (minb, dflt) in variants:
                    if SU != 4 and amp != "c128":
                        continue
                    out.append((f"amajor_{amp}_d{dt}", amp, SU, NT, minb, dt, dflt))
    return out


def direct_instances():
    out = []   # (group, amp, SU, NT, MINB, DT, is_default)
    for amp in ("c128", "c64", "r64"):
        for dt in range(1, DT_MAX + 1):
            for (SU, NT), variants in DIRECT.items():
                for (minb, dflt) in variants:
                    if not dflt and amp != "c128":
                        continue
                    out.append((f"amajor_{amp}_d{dt}", amp, SU, NT, minb, dt, dflt))
            for (SU, NT), variants in DIRECT_EXTRA.get(amp, {}).items():
                for (minb, dflt) in variants:
                    out.append((f"amajor_{amp}_d{dt}", amp, SU, NT, minb, dt, dflt))
    return out


def window_instances():
    out = []   # (group, amp, NT, MINB, DT, is_default)
    for amp in ("c128", "c64", "r64"):
        for dt in range(1, DT_MAX + 1):
            for (nt, minb, dflt) in WINDOW_CTAS:
                if not dflt and amp != "c128":
                    continue
                out.append((f"amajor_{amp}_d{dt}", amp, nt, minb, dt, dflt))
    return out


def write_if_changed(path, text):
    """keep timestamps stable so that `make` only rebuilds what really changed"""
    if os.path.exists(path) and open(path).read() == text:
        return
    with open(path, "w") as fh:
        fh.write(text)


def main():
    inst = instances()
    binst = block_instances()
    groups = {}
    for it in inst:
        groups.setdefault(it[0], []).append(it)
    bgroups = {}
    for it in binst:
        bgroups.setdefault(it[0], []).append(it)
    agroups = {}
    for it in amajor_instances():
        agroups.setdefault(it[0], []).append(it)
    dgroups = {}
    for it in direct_instances():
        dgroups.setdefault(it[0], []).append(it)
    egroups = {}
    for it in elem_instances():
        egroups.setdefault(it[0], []).append(it)
    wgroups = {}
    for it in window_instances():
        wgroups.setdefault(it[0], []).append(it)
    d = os.path.join(HERE, "instances")
    os.makedirs(d, exist_ok=True)
    import io
    wanted = set()
    names = []
    for g, items in sorted(groups.items()):
        names.append(g)
        fh = io.StringIO()
        fh.write("// generated by gen_instances.py - do not edit\n")
        fh.write('#include "../qkan_kernel.cuh"\n#include <vector>\nusing namespace qkan;\n')
        fh.write(f"void qkan_register_{g}(std::vector<KernelInfo>& reg) {{\n    KernelInfo k;\n")
        for (_, amp, L, NAT, NBT, T, NT, MINB, MODE, PREP, prio, vid, full) in items:
            A, R, _sz = AMPS[amp]
            fh.write(f"    k = make_info<{A}, {R}, {L}, {NAT}, {NBT}, {T}, {NT}, {MINB}, {MODE}, {PREP}, {full}>(); "
                     f"k.prio = {prio}; k.variant = {vid}; reg.push_back(k);\n")
        fh.write("}\n")
        wanted.add(f"inst_{g}.cu")
        write_if_changed(os.path.join(d, f"inst_{g}.cu"), fh.getvalue())
    bnames = []
    for g, items in sorted(bgroups.items()):
        bnames.append(g)
        fh = io.StringIO()
        fh.write("// generated by gen_instances.py - do not edit\n")
        fh.write('#include "../qkan_block.cuh"\n#include <vector>\nusing namespace qkan;\n')
        fh.write(f"void qkan_register_{g}(std::vector<BlockKernelInfo>& reg) {{\n")
        for (_, amp, U, SU, MODE, NT, MINB, DT, dflt) in items:
            A, R, _sz = AMPS[amp]
            fh.write(f"    reg.push_back(make_block_info<{A}, {R}, {U}, {SU}, {MODE}, {NT}, {MINB}, {DT}>({dflt}));\n")
        fh.write("}\n")
        wanted.add(f"inst_{g}.cu")
        write_if_changed(os.path.join(d, f"inst_{g}.cu"), fh.getvalue())
    for g, items in sorted(agroups.items()):
        bnames.append(g)
        fh = io.StringIO()
        fh.write("// generated by gen_instances.py - do not edit\n")
        fh.write('#include "../qkan_amajor.cuh"\n#include <vector>\nusing namespace qkan;\n')
        fh.write(f"void qkan_register_{g}(std::vector<BlockKernelInfo>& reg) {{\n")
        for (_, amp, SU, NT, MINB, DT, dflt) in items:
            A, R, _sz = AMPS[amp]
            fh.write(f"    reg.push_back(make_amajor_info<{A}, {R}, {SU}, {NT}, {MINB}, {DT}>({dflt}));\n")
        for (_, amp, SU, NT, MINB, DT, dflt) in dgroups.get(g, []):
            A, R, _sz = AMPS[amp]
            fh.write(f"    reg.push_back(make_direct_info<{A}, {R}, {SU}, {NT}, {MINB}, {DT}>({dflt}));\n")
        for (_, amp, SU, NT, MINB, DT, dflt) in egroups.get(g, []):
            A, R, _sz = AMPS[amp]
            fh.write(f"    reg.push_back(make_elem_info<{A}, {R}, {SU}, {NT}, {MINB}, {DT}>({dflt}));\n")
        for (_, amp, NT, MINB, DT, dflt) in wgroups.get(g, []):
            A, R, _sz = AMPS[amp]
            fh.write(f"    reg.push_back(make_amajor_window_info<{A}, {R}, {NT}, {MINB}, {DT}>({dflt}));\n")
        fh.write("}\n")
        wanted.add(f"inst_{g}.cu")
        write_if_changed(os.path.join(d, f"inst_{g}.cu"), fh.getvalue())
    for f in os.listdir(d):
        if f.startswith("inst_") and f.endswith(".cu") and f not in wanted:
            os.remove(os.path.join(d, f))
    if True:
        fh = io.StringIO()
        fh.write("// generated by gen_instances.py - do not edit\n#pragma once\n#include <vector>\n")
        fh.write("namespace qkan { struct KernelInfo; struct BlockKernelInfo; }\n")
        for g in names:
            fh.write(f"void qkan_register_{g}(std::vector<qkan::KernelInfo>& reg);\n")
        for g in bnames:
            fh.write(f"void qkan_register_{g}(std::vector<qkan::BlockKernelInfo>& reg);\n")
        fh.write("inline void qkan_register_all(std::vector<qkan::KernelInfo>& reg) {\n")
        for g in names:
            fh.write(f"    qkan_register_{g}(reg);\n")
        fh.write("}\n")
        fh.write("inline void qkan_register_all_block(std::vector<qkan::BlockKernelInfo>& reg) {\n")
        for g in bnames:
            fh.write(f"    qkan_register_{g}(reg);\n")
        fh.write("}\n")
        write_if_changed(os.path.join(HERE, "qkan_instances.h"), fh.getvalue())
    print(f"{len(inst)} staged + {len(binst)} generic block + {len(amajor_instances())} a-major + {len(window_instances())} window instances in {len(names) + len(bnames)} translation units")


if __name__ == "__main__":
    main()
