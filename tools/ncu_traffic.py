#!/usr/bin/env python
"""Record the DRAM traffic of a kernel launch from an `ncu --set full` capture into profiles/r02_traffic.json, keyed by the
workload and stamped with the hash of the kernel sources (bench.py quotes an entry only while that hash still matches).

    python tools/ncu_traffic.py gpurun_out/x.ncu-rep N K D B [dtype] [prep] [mode] [--name profiles-file-name]
"""
import csv
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from bench import csrc_sha  # noqa: E402


def main():
    argv = sys.argv[1:]
    name = None
    if "--name" in argv:
        i = argv.index("--name")
        name = argv[i + 1]
        argv = argv[:i] + argv[i + 2:]
    args = argv
    rep = args[0]
    N, K, D, B = (int(v) for v in args[1:5])
    dtype = args[5] if len(args) > 5 else "complex128"
    prep = args[6] if len(args) > 6 else "analytic"
    mode = args[7] if len(args) > 7 else "compat"
    name = name or os.path.basename(rep)
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, r = rows[0], rows[1], rows[-1]

    def val(metric):
        i = hdr.index(metric)
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1e-6, "msecond": 1e-3, "nsecond": 1e-9, "second": 1.0, "us": 1e-6, "ms": 1e-3, "ns": 1e-9, "s": 1.0}.get(units[i], 1.0)
        return float(r[i].replace(",", "")) * scale
    entry = {"file": name, "csrc_sha": csrc_sha(), "kernel": r[hdr.index("Kernel Name")],
             "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
             "gpu_time_s_under_ncu": val("gpu__time_duration.sum")}
    path = os.path.join(ROOT, "profiles", "r02_traffic.json")
    tab = json.load(open(path)) if os.path.exists(path) else {}
    tab[f"N{N}_K{K}_D{D}_B{B}_{dtype}_{prep}_{mode}"] = entry
    json.dump(tab, open(path, "w"), indent=1, sort_keys=True)
    print(json.dumps(entry))


if __name__ == "__main__":
    main()
