"""Drop-in host side of the reference's QKAN step / layer API, backed by the CUDA library.

Class names, constructor arguments, method names, return types and ValueError conditions
follow /root/reference/QKAN_Steps_original/{ChebyshevStep,MulStep,LCUStep,SUMStep,QKANLayer}.py
(cited per method).  Arithmetic never runs on the CPU here: every numeric method calls
libqkan_b200.so through ``_binding`` and raises if it is missing or no GPU is present.

Additions over the reference: ``forward`` accepts a batch ``[B, N]``
(NumPy or torch, host or CUDA) and returns ``[B, K]`` in the same container;
``dtype`` / ``mode`` / ``prep`` constructor keywords; ``return_amplitudes``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _binding as _b

try:  # torch is only needed for tensor inputs and for its CUDA stream / allocator
    import torch
except Exception:  # pragma: no cover
    torch = None

EPS_RANGE = 1e-8  # ChebyshevStep.py:26,40


def _is_torch(x) -> bool:
    return torch is not None and isinstance(x, torch.Tensor)


class _Engine:
    """Owns one qkan_layer handle (C ABI) and the host copy of the weight matrix."""

    def __init__(self, N: int, K: int, D: int, dtype="complex128", mode="compat", prep="analytic",
                 device: Optional[int] = None):
        if dtype not in _b.DTYPES:
            raise ValueError(f"dtype must be one of {list(_b.DTYPES)}")
        if mode not in _b.MODES:
            raise ValueError(f"mode must be one of {list(_b.MODES)}")
        if prep not in _b.PREPS:
            raise ValueError(f"prep must be one of {list(_b.PREPS)}")
        self.N, self.K, self.D = int(N), int(K), int(D)
        self.dtype, self.mode, self.prep = dtype, mode, prep
        self.device = device
        self._handle = None
        self._uploaded_version = -1

    # -- handle -----------------------------------------------------------------
    def handle(self):
        if self._handle is None:
            lib = _b.lib()
            if self.device is None:
                self.device = torch.cuda.current_device() if (torch is not None and torch.cuda.is_available()) else 0
            h = ctypes.c_void_p()
            _b.check(lib.qkan_layer_create(ctypes.byref(h), self.N, self.K, self.D, _b.DTYPES[self.dtype],
                                           _b.MODES[self.mode], _b.PREPS[self.prep], int(self.device)))
            self._handle = h
        return self._handle

    def close(self):
        if self._handle is not None:
            try:
                _b.lib().qkan_layer_destroy(self._handle)
            finally:
                self._handle = None

    def __del__(self):  # pragma: no cover
        try:
            self.close()
        except Exception:
            pass

    # -- weights ----------------------------------------------------------------
    def upload_host_weights(self, W: np.ndarray, version: int):
        if version == self._uploaded_version:
            return
        W = np.ascontiguousarray(W, dtype=np.float64)
        # host weights were validated by MulStep.set_weights; validate=0 keeps this asynchronous
        _b.check(_b.lib().qkan_layer_set_weights(self.handle(), W.ctypes.data, 0, 0, self._stream_ptr()))
        self._uploaded_version = version

    def upload_device_weights(self, W, validate: bool = True):
        # a rejected matrix (|w| > 1) leaves the previous tables in place; whatever happens, the host-side version
        # stamp no longer describes the device tables, so the next host upload is not skipped
        self._uploaded_version = -2
        _b.check(_b.lib().qkan_layer_set_weights(self.handle(), W.data_ptr(), 1, 1 if validate else 0,
                                                 self._stream_ptr()))

    def _stream_ptr(self):
        if torch is not None and torch.cuda.is_available():
            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        return None

    # -- calls ------------------------------------------------------------------
    def forward_host(self, x: np.ndarray, want_amps: bool, out: Optional[np.ndarray] = None):
        B = x.shape[0]
        if out is None:
            out = np.empty((B, self.K), dtype=np.float64)
        elif out.shape != (B, self.K) or out.dtype != np.float64 or not out.flags.c_contiguous:
            raise ValueError("out must be a C-contiguous float64 array of shape [B, K]")
        amps = None
        if want_amps:
            amps = np.empty((B, self.K), dtype=np.complex64 if self.dtype == "complex64" else np.complex128)
        _b.check(_b.lib().qkan_layer_forward_host(self.handle(), x.ctypes.data, B, out.ctypes.data,
                                                  amps.ctypes.data if want_amps else None))
        return out, amps

    def _check_device(self, t):
        """The engine owns tables on ONE device; tensors of another device would be handed to its kernel as raw pointers."""
        self.handle()
        if t.device.index != self.device:
            raise ValueError(f"tensor on cuda:{t.device.index} but this layer's engine lives on cuda:{self.device}; "
                             "create the layer with device=... or move the tensor")

    def forward_device(self, x, want_amps: bool):
        self._check_device(x)
        B = x.shape[0]
        out = torch.empty((B, self.K), dtype=torch.float64, device=x.device)
        amps = None
        if want_amps:
            amps = torch.empty((B, self.K), device=x.device,
                               dtype=torch.complex64 if self.dtype == "complex64" else torch.complex128)
        _b.check(_b.lib().qkan_layer_forward(self.handle(), x.data_ptr(), B, out.data_ptr(),
                                             amps.data_ptr() if want_amps else None, self._stream_ptr()))
        return out, amps

    def out_of_range(self) -> int:
        c = ctypes.c_uint64()
        _b.check(_b.lib().qkan_layer_out_of_range(self.handle(), ctypes.byref(c)))
        return int(c.value)

    def info(self) -> dict:
        ki = _b.KernelInfo()
        _b.check(_b.lib().qkan_layer_info(self.handle(), ctypes.byref(ki)))
        return ki.as_dict()

    def diagonals(self, x: np.ndarray, source: str = "closed_form"):
        """cheb [B,NK], weighted [B,D+1,NK], lcu [B,NK] computed on the GPU.  source="circuit": stage snapshots of the
        simulated circuit (post-selected block amplitudes); "closed_form": cos(D arccos x) like the reference."""
        if source not in ("closed_form", "circuit"):
            raise ValueError("source must be 'closed_form' or 'circuit'")
        if torch is None or not torch.cuda.is_available():
            raise RuntimeError("qkan_implementation_b200 needs a CUDA device (no CPU fallback)")
        dev = torch.device("cuda", self.device if self.device is not None else torch.cuda.current_device())
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(dev)
        B, NK = xd.shape[0], self.N * self.K
        cheb = torch.empty((B, NK), dtype=torch.float64, device=dev)
        wtd = torch.empty((B, self.D + 1, NK), dtype=torch.float64, device=dev)
        lcu = torch.empty((B, NK), dtype=torch.float64, device=dev)
        fn = _b.lib().qkan_layer_stage_snapshots if source == "circuit" else _b.lib().qkan_layer_diagonals
        _b.check(fn(self.handle(), xd.data_ptr(), B, cheb.data_ptr(), wtd.data_ptr(), lcu.data_ptr(), self._stream_ptr()))
        return cheb.cpu().numpy(), wtd.cpu().numpy(), lcu.cpu().numpy()


# ==============================================================================
class ChebyshevStep:
    """ChebyshevStep.py:8-65.  T_d(x) = cos(d arccos x), dilation by K."""

    def __init__(self, degree: int):
        if degree < 0:
            raise ValueError("Degree must be positive integer.")        # ChebyshevStep.py:14-15
        self.degree = degree
        self._cheb_engines = {}

    def _values(self, x: np.ndarray) -> np.ndarray:
        """T_degree of each entry, evaluated by the GPU diagonals kernel (N = len(x), K = 1, unit weight)."""
        x = np.atleast_1d(np.asarray(x, dtype=np.float64))
        n = x.shape[0]
        eng = self._cheb_engines.get(n)
        if eng is None:
            eng = _Engine(n, 1, self.degree)
            eng.upload_host_weights(np.ones((self.degree + 1, n)), 0)
            self._cheb_engines[n] = eng
        cheb, _, _ = eng.diagonals(x[None, :])
        return cheb[0]

    def apply_chebyshev(self, x: float) -> float:
        eps = EPS_RANGE
        if not np.all((x >= -1 - eps) & (x <= 1 + eps)):               # ChebyshevStep.py:26-27
            raise ValueError("Input value must be between -1 and 1.")
        v = self._values(np.atleast_1d(x))
        return v[0] if np.ndim(x) == 0 else v

    def transform_diagonal(self, x: np.ndarray) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        violations = x[~(-1 - EPS_RANGE <= x) | ~(x <= 1 + EPS_RANGE)]   # ChebyshevStep.py:46-49
        if len(violations) > 0:
            print(f"Values outside [-1,1] range: {violations[:5]}")
        return self._values(x)                                          # clip happens in the kernel (:52)

    def create_dilated_chebyshev(self, x: np.ndarray, K: int) -> np.ndarray:
        return np.diag(np.repeat(self.transform_diagonal(x), K))       # ChebyshevStep.py:62-65


class MulStep(ChebyshevStep):
    """MulStep.py:11-107.  Holds the [D+1, N*K] weights; weighted diagonal matrices."""

    def __init__(self, degree: int, num_weights: int):
        super().__init__(degree)
        self.num_weights = num_weights
        self._weights = np.zeros((degree + 1, num_weights))
        self._version = 0
        self._diag_engines = {}

    def set_weights(self, degree: int, weights: np.ndarray):
        if degree < 0 or degree > self.degree:                          # MulStep.py:32-33
            raise ValueError(f"Degree must be between 0 and {self.degree}")
        if len(weights) != self.num_weights:                            # MulStep.py:34-35
            raise ValueError(f"Expected {self.num_weights} weights, got {len(weights)}")
        if _is_torch(weights):
            weights = weights.detach().cpu().numpy()
        if not np.all(np.abs(weights) <= 1):                            # MulStep.py:36-37
            raise ValueError("Weight magnitudes must be <= 1 for unitarity")
        weights = np.asarray(weights, dtype=np.float64)
        if not np.array_equal(self._weights[degree], weights):          # unchanged rows keep the device tables valid
            self._weights[degree] = weights
            self._version += 1

    def _engine_for(self, N: int, K: int) -> _Engine:
        eng = self._diag_engines.get((N, K))
        if eng is None:
            eng = _Engine(N, K, self.degree)
            self._diag_engines[(N, K)] = eng
        eng.upload_host_weights(self._weights, self._version)
        return eng

    def _check_size(self, x, K):
        N = len(x)
        expected = N * K
        if self.num_weights != expected:                                # MulStep.py:62-66
            raise ValueError(f"Weight vector size {self.num_weights} does not match "
                             f"expected size {expected} = {N}*{K}")
        return N

    def get_weighted_polynomial_matrix(self, x: np.ndarray, K: int, degree: int) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        violations = x[~(-1 - EPS_RANGE <= x) | ~(x <= 1 + EPS_RANGE)]
        if len(violations) > 0:
            print(f"Values outside [-1,1] range: {violations[:5]}")
        N = self._check_size(x, K)
        _, wtd, _ = self._engine_for(N, K).diagonals(x[None, :])
        return np.diag(wtd[0, degree])                                  # MulStep.py:69-72

    def create_weighted_chebyshev(self, x: np.ndarray, K: int, degree: int):
        from .fable import fable
        return fable(self.get_weighted_polynomial_matrix(x, K, degree), 0)   # MulStep.py:107


class LCUStep:
    """LCUStep.py:10-60."""

    def __init__(self, max_degree: int):
        self.max_degree = max_degree

    def get_combined_matrix(self, x: np.ndarray, mul_step: MulStep, K: int) -> np.ndarray:
        x = np.asarray(x, dtype=np.float64)
        N = mul_step._check_size(x, K)
        violations = x[~(-1 - EPS_RANGE <= x) | ~(x <= 1 + EPS_RANGE)]
        if len(violations) > 0:
            print(f"Values outside [-1,1] range: {violations[:5]}")
        _, _, lcu = mul_step._engine_for(N, K).diagonals(x[None, :])
        return np.diag(lcu[0])                                          # LCUStep.py:32-37

    def combine_weighted_polynomials(self, x: np.ndarray, mul_step: MulStep, K: int):
        from .fable import fable
        return fable(self.get_combined_matrix(x, mul_step, K), 0)      # LCUStep.py:60


class SUMStep:
    """SUMStep.py:10-31."""

    def __init__(self):
        pass

    def apply_sum(self, matrix: np.ndarray, N: int, K: int):
        from .fable import fable
        diag_elements = np.diag(matrix).reshape(N, K, order="F")       # SUMStep.py:28
        summed = np.sum(diag_elements, axis=0) / N
        return fable(np.diag(summed), 0)                               # SUMStep.py:30-31


class QKANLayer:
    """QKANLayer.py:12-135, batched and GPU-resident."""

    def __init__(self, N: int, K: int, max_degree: int, *, dtype: str = "complex128", mode: str = "compat",
                 prep: str = "analytic", device: Optional[int] = None):
        self.N = N
        self.K = K
        self.max_degree = max_degree
        self.cheb_step = ChebyshevStep(max_degree)
        self.mul_step = MulStep(max_degree, N * K)
        self.lcu_step = LCUStep(max_degree)
        self.sum_step = SUMStep()
        self._engine = _Engine(N, K, max_degree, dtype=dtype, mode=mode, prep=prep, device=device)
        self.dtype, self.mode, self.prep = dtype, mode, prep

    # ------------------------------------------------------------------ weights
    def _set_weights(self, weights, check_len: bool = False):
        """QKANLayer.py:124-125 (fast path) / :46-49 (verbose path)."""
        if _is_torch(weights) and weights.is_cuda:
            if weights.dim() != 2 or weights.shape[0] > self.max_degree + 1:
                raise ValueError(f"Degree must be between 0 and {self.max_degree}")
            if weights.shape[1] != self.N * self.K:
                raise ValueError(f"Expected {self.N * self.K} weights, got {weights.shape[1]}")
            self._engine._check_device(weights)
            # the same tensor, unmodified since its last upload: tables, host mirror and validation are all current
            stamp = (weights.data_ptr(), weights._version, tuple(weights.shape), weights.dtype)
            if getattr(self, "_device_stamp", None) == stamp and self._engine._uploaded_version == self.mul_step._version:
                return
            if weights.shape[0] != self.max_degree + 1 or weights.dtype != torch.float64 or not weights.is_contiguous():
                full = torch.as_tensor(self.mul_step._weights, device=weights.device).clone()
                full[: weights.shape[0]] = weights.to(torch.float64)
                weights = full
            self._device_stamp = None
            try:
                self._engine.upload_device_weights(weights, validate=True)
            except _b.QkanError as e:
                if e.code == _b.ERR_WEIGHT_RANGE:
                    raise ValueError("Weight magnitudes must be <= 1 for unitarity") from None
                raise
            self._device_weights = weights
            self.mul_step._weights[:] = weights.detach().cpu().numpy()
            self.mul_step._version += 1
            self._engine._uploaded_version = self.mul_step._version
            self._device_stamp = stamp
            return
        for degree, w in enumerate(weights):
            if check_len and len(w) != self.N * self.K:
                raise ValueError(f"Expected weight dimension {self.N * self.K}")
            self.mul_step.set_weights(degree, w)
        self._engine.upload_host_weights(self.mul_step._weights, self.mul_step._version)

    # ------------------------------------------------------------------ forward
    def forward(self, x, weights, verbose: bool = False, return_amplitudes: bool = False,
                check_range: Optional[bool] = None, out=None):
        """x: [N] -> [K] (like the reference) or [B, N] -> [B, K].  NumPy in -> NumPy out,
        torch in -> torch out on the same device."""
        if verbose:
            return self._forward_verbose(x, weights)
        self._set_weights(weights)
        is_t = _is_torch(x)
        if not is_t:
            x = np.asarray(x, dtype=np.float64)
        if x.ndim not in (1, 2):
            raise ValueError(f"x must be [N] or [B, N], got shape {tuple(x.shape)}")
        n_in = x.shape[-1]
        if self.mul_step.num_weights != n_in * self.K:                  # MulStep.py:62-66
            raise ValueError(f"Weight vector size {self.mul_step.num_weights} does not match "
                             f"expected size {n_in * self.K} = {n_in}*{self.K}")
        single = (x.ndim == 1)
        if is_t and x.is_cuda:
            xd = x.to(torch.float64).contiguous()
            if single:
                xd = xd[None, :]
            out, amps = self._engine.forward_device(xd, return_amplitudes)
            if check_range:
                self._warn_range(None)
        else:
            xh = x.detach().numpy() if is_t else np.asarray(x)
            xh = np.ascontiguousarray(xh, dtype=np.float64)
            if single:
                xh = xh[None, :]
            out, amps = self._engine.forward_host(xh, return_amplitudes, out=out)
            if check_range is None or check_range:
                self._warn_range(xh)
            if is_t:
                out = torch.from_numpy(out)
                amps = torch.from_numpy(amps) if amps is not None else None
        if single:
            out = out[0]
            amps = amps[0] if amps is not None else None
        return (out, amps) if return_amplitudes else out

    __call__ = forward

    def _warn_range(self, xh):
        n = self._engine.out_of_range()
        if n:
            if xh is not None:
                v = xh[~(-1 - EPS_RANGE <= xh) | ~(xh <= 1 + EPS_RANGE)]
                print(f"Values outside [-1,1] range: {v[:5]}")          # ChebyshevStep.py:48-49
            else:
                print(f"Values outside [-1,1] range: {n} entries (clipped)")

    def out_of_range_count(self) -> int:
        return self._engine.out_of_range()

    def kernel_info(self) -> dict:
        return self._engine.info()

    # ------------------------------------------------------ intermediate matrices
    def get_intermediate_matrices(self, x, weights, source: str = "circuit") -> dict:
        """QKANLayer.py:30-75: dense intermediates of one sample.  The diagonals are stage snapshots of the simulated
        circuit (source="circuit": the post-selected block amplitudes after CHEB, after SELECT and after the degree
        sum) or the reference's closed form evaluated on the GPU (source="closed_form")."""
        if len(x) != self.N:
            raise ValueError(f"Expected input dimension {self.N}, got {len(x)}")
        if len(weights) != self.max_degree + 1:
            raise ValueError(f"Expected {self.max_degree + 1} weight vectors")
        self._set_weights(weights, check_len=True)
        x = np.asarray(x, dtype=np.float64)
        violations = x[~(-1 - EPS_RANGE <= x) | ~(x <= 1 + EPS_RANGE)]
        if len(violations) > 0:
            print(f"Values outside [-1,1] range: {violations[:5]}")
        cheb, wtd, lcu = self._engine.diagonals(x[None, :], source)
        D = self.max_degree
        results = {"input": x}
        results["cheb"] = {d: np.diag(cheb[0]) for d in range(D + 1)}          # degree quirk, :54-57
        results["weighted"] = {d: np.diag(wtd[0, d]) for d in range(D + 1)}
        results["lcu"] = np.diag(lcu[0])
        results["reshaped"] = lcu[0].reshape(self.N, self.K, order="F")        # QKANLayer.py:70
        out = self._engine.forward_host(x[None, :], False)[0][0]
        results["final"] = out
        return results

    def get_intermediate_diagonals(self, x, source: str = "circuit") -> dict:
        """Batched variant: x [B, N] -> dict of diagonals cheb [B,NK], weighted [B,D+1,NK], lcu [B,NK],
        reshaped [B,N,K], final [B,K] (weights as last set); `source` as in get_intermediate_matrices."""
        x = np.atleast_2d(np.asarray(x, dtype=np.float64))
        self._engine.upload_host_weights(self.mul_step._weights, self.mul_step._version)
        cheb, wtd, lcu = self._engine.diagonals(x, source)
        reshaped = lcu.reshape(-1, self.K, self.N).transpose(0, 2, 1)
        final = self._engine.forward_host(np.ascontiguousarray(x), False)[0]
        return {"cheb": cheb, "weighted": wtd, "lcu": lcu, "reshaped": reshaped, "final": final}

    def _forward_verbose(self, x, weights):
        """QKANLayer.py:90-120."""
        m = self.get_intermediate_matrices(x, weights)
        print("\nQKAN Layer Forward Pass:")
        print(f"Input x: {m['input']}")
        print("\nStep 1-2 (DILATE + CHEB):")
        for d, mat in m["cheb"].items():
            print(f"Chebyshev matrix degree {d}:")
            print(mat)
            print("Matrix diagonal:", np.diag(mat))
        print("\nStep 3 (MUL):")
        for d, mat in m["weighted"].items():
            print(f"Weighted matrix degree {d}:")
            print(mat)
            print("Matrix diagonal:", np.diag(mat))
        print("\nStep 4 (LCU):")
        print("Combined matrix:")
        print(m["lcu"])
        print("Matrix diagonal:", np.diag(m["lcu"]))
        print("\nStep 5 (SUM):")
        print("Reshaped (NxK):")
        print(m["reshaped"])
        print("Final output (summed over inputs):")
        print(m["final"])
        return m["final"]
