#!/usr/bin/env python
"""Per-loop instruction and stall-sample shares of a kernel from an .ncu-rep (SASS source page).
    python tools/ncu_sass.py file.ncu-rep [--dump]
Splits the kernel at its backward branches (innermost loops first), and for every region prints the warp instructions
executed, how many of them are FP-pipe arithmetic, and the share of warp-stall samples."""
import csv
import io
import re
import subprocess
import sys

FP = re.compile(r"\bD(FMA|MUL|ADD|SETP|MNMX)|MUFU.*64|\bF(FMA|MUL|ADD)\b")


def main():
    rep = sys.argv[1]
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = rows[1]
    ia, isrc, ismp, iex = hdr.index("Address"), hdr.index("Source"), hdr.index("Warp Stall Sampling (All Samples)"), hdr.index("Instructions Executed")
    body = [r for r in rows[2:] if len(r) > iex and r[ia].startswith("0x")]
    base = int(body[0][ia], 16)
    ins = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[ismp] or 0), int(r[iex] or 0)) for r in body]
    tot_s = sum(i[2] for i in ins) or 1
    tot_e = sum(i[3] for i in ins) or 1
    fp_e = sum(i[3] for i in ins if FP.search(i[1]))
    print(f"{len(ins)} SASS instructions; executed {tot_e} warp-instr, FP-pipe {fp_e} ({100.0 * fp_e / tot_e:.1f} %); samples {tot_s}")
    if "--dump" in sys.argv:
        for a, s, sm, ex in ins:
            print(f"{a:6x} {ex:10d} {100.0 * sm / tot_s:6.2f}%  {s}")
        return
    loops = []
    for a, s, _, _ in ins:
        m = re.search(r"BRA.*0x([0-9a-f]+)", s)
        if m:
            t = int(m.group(1), 16) - base if int(m.group(1), 16) >= base else int(m.group(1), 16)
            if t < a:
                loops.append((t, a))
    loops.sort(key=lambda l: l[1] - l[0])
    owner = {}
    for li, (t, a) in enumerate(loops):
        for ad, *_ in ins:
            if t <= ad <= a and ad not in owner:
                owner[ad] = li
    agg = {}
    for ad, s, sm, ex in ins:
        k = owner.get(ad, -1)
        g = agg.setdefault(k, [0, 0, 0, 0])
        g[0] += ex
        g[1] += ex if FP.search(s) else 0
        g[2] += sm
        g[3] += 1
    for k in sorted(agg, key=lambda k: -agg[k][0]):
        ex, fp, sm, n = agg[k]
        name = "outside loops" if k < 0 else f"loop {loops[k][0]:#x}..{loops[k][1]:#x}"
        print(f"  {name:28s} {n:5d} instr  executed {100.0 * ex / tot_e:6.2f} %  FP {100.0 * fp / max(ex, 1):5.1f} %  samples {100.0 * sm / tot_s:6.2f} %")


if __name__ == "__main__":
    main()
