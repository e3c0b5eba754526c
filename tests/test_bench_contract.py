"""bench.py's driver contract, checked where no GPU is needed: the reference arm (`--impl reference`, CPU port of the reference
algorithm) prints ONE JSON line with the contract keys, alone and under torchrun (rank 0 prints, the others exit 0); the
traffic table quoted by the GPU arm matches the kernel sources in the tree."""
import json
import os
import subprocess
import sys

from conftest import ROOT

KEYS = ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "vs_baseline",
        "dtype", "data", "config", "impl", "cpu_baseline", "e2e")


def _check(line, n_gpus):
    d = json.loads(line)
    for k in KEYS:
        assert k in d, k
    assert d["impl"] == "reference" and d["higher_is_better"] is True and d["vs_baseline"] is None
    assert d["n_gpus"] == n_gpus and d["value"] > 0 and d["unit"] == "samples/s"
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"]["value"] == d["value"] and d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and "model" not in d["config"]


def test_reference_arm_prints_one_contract_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--cpu-samples", "200"], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 1)


def test_reference_arm_under_torchrun_rank0_only():
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
                          "--master-port", "29931", os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
                          "--warmup", "1", "--cpu-samples", "200"], capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    _check(lines[0], 2)


def test_committed_traffic_matches_the_kernel_sources():
    """`roofline.traffic` is quoted from profiles/r02_traffic.json only while the hash of the kernel headers matches: the
    committed table must belong to the tree it is committed with (a stale table silently turns the key into null)."""
    sys.path.insert(0, ROOT)
    import bench
    tab = json.load(open(os.path.join(ROOT, "profiles", "r02_traffic.json")))
    sha = bench.csrc_sha()
    for key in ("N4_K4_D3_B1000000_complex128_analytic_compat", "N16_K16_D8_B1000000_complex128_analytic_compat",
                "N784_K10_D5_B100000_complex128_analytic_compat"):
        assert tab[key]["csrc_sha"] == sha, (key, tab[key]["csrc_sha"], sha)
        traffic, src = bench.measured_traffic(key)
        assert traffic and traffic > 0 and "profiles/" in src
