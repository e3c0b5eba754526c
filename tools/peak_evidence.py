#!/usr/bin/env python
"""Run the two FP64 peak microbenchmarks that the roofline denominators come from (qkan_measure_fma_peak: independent DFMA
chains; qkan_measure_dmma_peak: independent mma.sync.m8n8k4.f64 chains).  Under ncu this gives their pipe utilisation:
    ncu --metrics sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,gpu__time_duration.sum,sm__cycles_elapsed.max \
        --clock-control none -k regex:peak_kernel python tools/peak_evidence.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from qkan_implementation_b200 import _binding as b  # noqa: E402

print(f"DFMA peak  {b.measure_fma_peak(0, True):.2f} TFLOP/s (2 flops per lane-FMA)")
print(f"FFMA peak  {b.measure_fma_peak(0, False):.2f} TFLOP/s")
print(f"DMMA peak  {b.measure_dmma_peak(0):.2f} TFLOP/s (mma.sync.m8n8k4.f64, 512 flops per warp instruction)")
