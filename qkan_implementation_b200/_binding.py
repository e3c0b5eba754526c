"""ctypes binding of libqkan_b200.so (C ABI declared in include/qkan_b200.h).

There is no CPU fallback: if the CUDA library is missing or cannot be loaded every
entry point raises.  Build it with ``python -c "import __graft_entry__ as g; g.build()"``
or ``make -C qkan_implementation_b200/csrc -j8``.
"""
from __future__ import annotations

import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("QKAN_B200_LIB", os.path.join(_HERE, "libqkan_b200.so"))   # override = tuning aid

QKAN_OK = 0
ERR_BAD_SHAPE, ERR_UNSUPPORTED, ERR_WEIGHT_RANGE, ERR_CUDA, ERR_NO_WEIGHTS = -1, -2, -3, -4, -5
DTYPES = {"complex128": 0, "complex64": 1, "real64": 2}
MODES = {"compat": 0, "paper": 1}
PREPS = {"analytic": 1, "gates": 0}

EXPORTS = [
    "qkan_layer_create", "qkan_layer_destroy", "qkan_layer_set_weights", "qkan_layer_forward",
    "qkan_layer_forward_host", "qkan_plan_host_chunks", "qkan_layer_forward_peers", "qkan_layer_forward_multicast", "qkan_layer_out_of_range", "qkan_layer_info", "qkan_layer_diagonals", "qkan_layer_stage_snapshots",
    "qkan_forward", "qkan_measure_fma_peak", "qkan_last_error", "qkan_version", "qkan_simulate_circuit",
    "qkan_set_last_error",
    "qkan_measure_dmma_peak", "qkan_cheb_gram_workspace", "qkan_cheb_gram", "qkan_cheb_residuals_ctas", "qkan_cheb_residuals", "qkan_cheb_features",
]


class KernelInfo(ctypes.Structure):
    _fields_ = [(n, ctypes.c_int) for n in
                ("engine", "n_a", "n_b", "l", "qubits",
                 "blocks", "unroll", "samples_per_lane", "lanes_per_sample", "lanes_per_row", "rows_in_parallel", "passes", "row_steps",
                 "tile_qubits", "tile_na", "tile_nb", "local_qubits", "stages", "sectors_total", "sectors_run",
                 "threads_per_cta", "min_ctas_per_sm", "samples_per_cta", "grid", "smem_bytes",
                 "passes_survey", "passes_exec", "scaled_rotations", "input_window", "element_owner", "degree_factored", "cheb_elements", "direct_rows")] + \
               [(n, ctypes.c_double) for n in
                ("flops_survey", "flops_per_block_basis", "flops_exec", "fp_inst_exec", "layout_efficiency", "io_bytes")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class QkanError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"qkan_b200 error {code}: {msg}")
        self.code = code
        self.message = msg


_lib = None


def lib():
    """Load the shared library once; raise (loudly) if it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} not found: the CUDA extension is not built. There is no CPU fallback - "
            "run `make -C qkan_implementation_b200/csrc -j8` (needs nvcc, targets sm_100a).")
    L = ctypes.CDLL(LIB_PATH)
    vp, i32, i64 = ctypes.c_void_p, ctypes.c_int, ctypes.c_int64
    L.qkan_layer_create.argtypes = [ctypes.POINTER(vp), i32, i32, i32, i32, i32, i32, i32]
    L.qkan_layer_destroy.argtypes = [vp]
    L.qkan_layer_destroy.restype = None
    L.qkan_layer_set_weights.argtypes = [vp, vp, i32, i32, vp]
    L.qkan_layer_forward.argtypes = [vp, vp, i64, vp, vp, vp]
    L.qkan_layer_forward_host.argtypes = [vp, vp, i64, vp, vp]
    L.qkan_layer_forward_peers.argtypes = [vp, vp, i64, ctypes.POINTER(vp), i32, i64, vp]
    L.qkan_layer_forward_multicast.argtypes = [vp, vp, i64, vp, i64, vp]
    L.qkan_layer_out_of_range.argtypes = [vp, ctypes.POINTER(ctypes.c_uint64)]
    L.qkan_layer_info.argtypes = [vp, ctypes.POINTER(KernelInfo)]
    L.qkan_layer_diagonals.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.qkan_layer_stage_snapshots.argtypes = [vp, vp, i64, vp, vp, vp, vp]
    L.qkan_forward.argtypes = [vp, vp, vp, i64, i32, i32, i32, i32, i32, vp, vp]
    L.qkan_simulate_circuit.argtypes = [vp, vp, i32, i32, vp, i64, vp, vp]
    L.qkan_measure_fma_peak.argtypes = [i32, i32, ctypes.POINTER(ctypes.c_double)]
    L.qkan_measure_dmma_peak.argtypes = [i32, ctypes.POINTER(ctypes.c_double)]
    L.qkan_cheb_gram_workspace.argtypes = [i64, i32, i32, ctypes.POINTER(i64), ctypes.POINTER(i32)]
    L.qkan_cheb_gram.argtypes = [vp, vp, i64, i32, i32, vp, vp, i64, vp]
    L.qkan_cheb_residuals_ctas.argtypes = [ctypes.POINTER(i32)]
    L.qkan_cheb_residuals.argtypes = [vp, vp, vp, i64, i32, i32, vp, ctypes.c_double, vp, vp, vp, vp]
    L.qkan_cheb_features.argtypes = [vp, i64, i32, i32, vp, vp]
    L.qkan_last_error.restype = ctypes.c_char_p
    L.qkan_version.argtypes = [ctypes.POINTER(i32)] * 3
    L.qkan_version.restype = None
    _lib = L
    return L


def check(rc: int):
    if rc != QKAN_OK:
        raise QkanError(rc, lib().qkan_last_error().decode())


def measure_dmma_peak(device: int = 0) -> float:
    v = ctypes.c_double()
    check(lib().qkan_measure_dmma_peak(device, ctypes.byref(v)))
    return v.value


def measure_fma_peak(device: int = 0, fp64: bool = True) -> float:
    v = ctypes.c_double()
    check(lib().qkan_measure_fma_peak(device, 1 if fp64 else 0, ctypes.byref(v)))
    return v.value
