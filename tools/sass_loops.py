#!/usr/bin/env python
"""Instruction mix of the loops in a kernel's SASS (cuobjdump -sass): for every backward branch prints the loop body size
and how many of its instructions run on the FP64 pipe.   python tools/sass_loops.py file.o [substring of the mangled name]"""
import re
import subprocess
import sys


def main():
    obj, pat = sys.argv[1], (sys.argv[2] if len(sys.argv) > 2 else "")
    txt = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True, check=True).stdout
    funcs = re.split(r"\n\s*Function : ", txt)[1:]
    for f in funcs:
        name = f.split("\n", 1)[0].strip()
        if pat not in name:
            continue
        ops = []
        for l in f.split("\n"):
            m = re.search(r"/\*([0-9a-f]{4,6})\*/\s+(.*?);", l)
            if m:
                ops.append((int(m.group(1), 16), m.group(2).strip()))
        fp_re = re.compile(r"\bD(FMA|MUL|ADD|SETP|MNMX)|MUFU.*64|\bF(FMA|MUL|ADD)\b")
        tot_fp = sum(1 for _, o in ops if fp_re.search(o))
        print(f"{name}\n  {len(ops)} instructions, {tot_fp} FP-pipe")
        for addr, op in ops:
            m = re.search(r"BRA.*0x([0-9a-f]+)", op)
            if m and int(m.group(1), 16) < addr:
                tgt = int(m.group(1), 16)
                body = [o for a, o in ops if tgt <= a <= addr]
                fp = sum(1 for o in body if fp_re.search(o))
                kinds = {}
                for o in body:
                    if fp_re.search(o):
                        continue
                    k = re.sub(r"^@!?U?P\d\s+", "", o).split()[0].split(".")[0]
                    kinds[k] = kinds.get(k, 0) + 1
                if len(body) >= 24:
                    print(f"  loop {tgt:#x}..{addr:#x}: {len(body)} instr, FP {fp}, other {len(body) - fp}  {dict(sorted(kinds.items(), key=lambda kv: -kv[1]))}")


if __name__ == "__main__":
    main()
