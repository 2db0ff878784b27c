"""ctypes binding of the C ABI in include/pinn_b200.h.

This is the Python stand-in for the Julia shim of INTEGRATION.md: it builds a
``pinn_problem_desc`` from plain Python data and calls the library.  There is no CPU
fallback: if the shared library is missing or no CUDA device is present the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field
from typing import List, Optional, Sequence

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PINN_B200_LIB") or os.path.join(HERE, "lib", "libpinn_b200.so")

ABI_VERSION = 2
MAX_IN = 8
MAX_TERMS = 32

F32, F64 = 0, 1
MODE_FFMA, MODE_TC_BF16, MODE_TC_SPLIT = 0, 1, 2
ACT = {"identity": 0, "tanh": 1, "sigmoid": 2, "sin": 3, "softplus": 4, "swish": 5}
OP = {
    "const": 0, "coord": 1, "tap": 2, "param": 3, "add": 4, "sub": 5, "mul": 6, "div": 7, "neg": 8,
    "pow": 9, "powi": 10, "sin": 11, "cos": 12, "exp": 13, "log": 14, "tanh": 15, "sqrt": 16, "abs": 17,
}
REDUCE_MEAN, REDUCE_WSUM = 0, 1

# symbols include/pinn_b200.h declares; tests check that the library exports every one
EXPORTS = [
    "pinn_create", "pinn_destroy", "pinn_last_error", "pinn_abi_version", "pinn_set_points",
    "pinn_set_points_host", "pinn_set_global_count", "pinn_loss_grad", "pinn_loss_grad_host",
    "pinn_term_residual", "pinn_term_residual_host", "pinn_comm_unique_id", "pinn_comm_init",
    "pinn_launch_count", "pinn_set_timing", "pinn_last_kernel_ms", "pinn_workspace_bytes",
    "pinn_flops_per_eval", "pinn_adam_begin", "pinn_adam_iterate", "pinn_adam_theta",
    "pinn_term_grad_stats", "pinn_term_grad_stats_host", "pinn_set_sampler", "pinn_resample", "pinn_get_points_host",
    "pinn_comm_info", "pinn_set_sampler_ex",
]


class EngineError(RuntimeError):
    """Raised for any nonzero return code of the C ABI (message from pinn_last_error)."""


class _Instr(C.Structure):
    _fields_ = [("op", C.c_int32), ("a", C.c_int32), ("b", C.c_int32), ("_pad", C.c_int32), ("imm", C.c_double)]


class _NetDesc(C.Structure):
    _fields_ = [("n_layers", C.c_int32), ("dims", C.POINTER(C.c_int32)), ("acts", C.POINTER(C.c_int32)),
                ("theta_offset", C.c_int64)]


class _TapDesc(C.Structure):
    _fields_ = [("net", C.c_int32), ("out", C.c_int32), ("order", C.c_int32), ("dir", C.c_int32 * 4)]


class _TermDesc(C.Structure):
    _fields_ = [("dim", C.c_int32), ("n_taps", C.c_int32), ("taps", C.POINTER(_TapDesc)),
                ("net_rows", C.POINTER(C.c_int32)), ("n_instr", C.c_int32), ("prog", C.POINTER(_Instr)),
                ("reduction", C.c_int32), ("scale", C.c_double)]


class _ProblemDesc(C.Structure):
    _fields_ = [("abi_version", C.c_int32), ("dtype", C.c_int32), ("mode", C.c_int32), ("device", C.c_int32),
                ("n_nets", C.c_int32), ("nets", C.POINTER(_NetDesc)), ("n_terms", C.c_int32),
                ("terms", C.POINTER(_TermDesc)), ("n_params", C.c_int32), ("param_offset", C.c_int64),
                ("n_theta", C.c_int64)]


# ---- plain-data problem description (what the Julia shim would assemble) ---------------------
@dataclass
class NetSpec:
    dims: Sequence[int]                 # in, hidden..., out
    acts: Sequence[str]                 # one per Dense layer
    theta_offset: int = 0

    @property
    def n_params(self) -> int:
        return sum(self.dims[i] * self.dims[i + 1] + self.dims[i + 1] for i in range(len(self.dims) - 1))


@dataclass
class TapSpec:
    net: int
    order: int = 0
    dirs: Sequence[int] = ()
    out: int = 0


@dataclass
class TermSpec:
    dim: int
    taps: List[TapSpec]
    prog: List[tuple]                   # (opname, a, b, imm)
    net_rows: Optional[List[List[int]]] = None   # per network: point row feeding input j
    reduction: int = REDUCE_MEAN
    scale: float = 1.0


@dataclass
class ProblemSpec:
    nets: List[NetSpec]
    terms: List[TermSpec]
    n_params: int = 0
    param_offset: int = 0
    n_theta: int = 0
    dtype: str = "float32"
    mode: int = MODE_FFMA
    device: int = 0
    _keep: list = field(default_factory=list, repr=False)


_lib = None


def load_library():
    """Load libpinn_b200.so and declare prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise EngineError(
            "libpinn_b200.so is not built (%s); run `python __graft_entry__.py build` -- "
            "this engine has no CPU fallback" % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    lib.pinn_create.argtypes = [C.POINTER(_ProblemDesc), C.POINTER(vp)]
    lib.pinn_create.restype = C.c_int
    lib.pinn_destroy.argtypes = [vp]
    lib.pinn_destroy.restype = C.c_int
    lib.pinn_last_error.argtypes = []
    lib.pinn_last_error.restype = C.c_char_p
    lib.pinn_abi_version.argtypes = []
    lib.pinn_abi_version.restype = C.c_int
    lib.pinn_set_points.argtypes = [vp, i32, vp, i64, vp]
    lib.pinn_set_points.restype = C.c_int
    lib.pinn_set_points_host.argtypes = [vp, i32, vp, i64, vp, vp]
    lib.pinn_set_points_host.restype = C.c_int
    lib.pinn_set_global_count.argtypes = [vp, i32, i64]
    lib.pinn_set_global_count.restype = C.c_int
    lib.pinn_loss_grad.argtypes = [vp, vp, C.POINTER(dbl), vp, vp, vp, vp]
    lib.pinn_loss_grad.restype = C.c_int
    lib.pinn_loss_grad_host.argtypes = [vp, vp, C.POINTER(dbl), vp, vp, vp]
    lib.pinn_loss_grad_host.restype = C.c_int
    lib.pinn_set_sampler.argtypes = [vp, i32, i64, C.POINTER(dbl), C.POINTER(dbl), C.c_uint64, vp]
    lib.pinn_set_sampler.restype = C.c_int
    lib.pinn_set_sampler_ex.argtypes = [vp, i32, i32, i64, C.POINTER(dbl), C.POINTER(dbl), C.c_uint64, vp]
    lib.pinn_set_sampler_ex.restype = C.c_int
    lib.pinn_resample.argtypes = [vp, vp]
    lib.pinn_resample.restype = C.c_int
    lib.pinn_get_points_host.argtypes = [vp, i32, vp]
    lib.pinn_get_points_host.restype = C.c_int
    lib.pinn_term_grad_stats.argtypes = [vp, i32, vp, C.POINTER(dbl), C.POINTER(dbl), vp]
    lib.pinn_term_grad_stats.restype = C.c_int
    lib.pinn_term_grad_stats_host.argtypes = [vp, i32, vp, C.POINTER(dbl), C.POINTER(dbl)]
    lib.pinn_term_grad_stats_host.restype = C.c_int
    lib.pinn_term_residual.argtypes = [vp, i32, vp, vp, vp]
    lib.pinn_term_residual.restype = C.c_int
    lib.pinn_term_residual_host.argtypes = [vp, i32, vp, vp]
    lib.pinn_term_residual_host.restype = C.c_int
    lib.pinn_comm_unique_id.argtypes = [vp]
    lib.pinn_comm_unique_id.restype = C.c_int
    lib.pinn_comm_init.argtypes = [vp, vp, i32, i32]
    lib.pinn_comm_init.restype = C.c_int
    lib.pinn_comm_info.argtypes = [vp, C.POINTER(i32)]
    lib.pinn_comm_info.restype = C.c_char_p
    lib.pinn_launch_count.argtypes = [vp]
    lib.pinn_launch_count.restype = i64
    lib.pinn_set_timing.argtypes = [vp, i32]
    lib.pinn_set_timing.restype = C.c_int
    lib.pinn_last_kernel_ms.argtypes = [vp]
    lib.pinn_last_kernel_ms.restype = dbl
    lib.pinn_workspace_bytes.argtypes = [vp]
    lib.pinn_workspace_bytes.restype = i64
    lib.pinn_flops_per_eval.argtypes = [vp]
    lib.pinn_flops_per_eval.restype = dbl
    lib.pinn_adam_begin.argtypes = [vp, vp, dbl, dbl, dbl, dbl]
    lib.pinn_adam_begin.restype = C.c_int
    lib.pinn_adam_iterate.argtypes = [vp, i32, C.POINTER(dbl), vp, vp]
    lib.pinn_adam_iterate.restype = C.c_int
    lib.pinn_adam_theta.argtypes = [vp, vp]
    lib.pinn_adam_theta.restype = C.c_int
    _lib = lib
    return lib


def _check(rc: int):
    if rc != 0:
        raise EngineError(load_library().pinn_last_error().decode("utf-8", "replace"))


def _np_dtype(dtype: str):
    return np.float64 if dtype in ("float64", "f64") else np.float32


def _ptr(x) -> C.c_void_p:
    """Device/host pointer of a numpy array, torch tensor, int address or None."""
    if x is None:
        return C.c_void_p(0)
    if isinstance(x, int):
        return C.c_void_p(x)
    if isinstance(x, np.ndarray):
        return C.c_void_p(x.ctypes.data)
    if hasattr(x, "data_ptr"):
        return C.c_void_p(x.data_ptr())
    raise TypeError("cannot take a pointer of %r" % type(x))


def build_desc(spec: ProblemSpec) -> _ProblemDesc:
    """Marshal a ProblemSpec into the C descriptor (buffers are kept alive on the spec)."""
    keep = spec._keep
    keep.clear()
    nets = (_NetDesc * len(spec.nets))()
    for k, n in enumerate(spec.nets):
        dims = (C.c_int32 * len(n.dims))(*[int(v) for v in n.dims])
        acts = (C.c_int32 * len(n.acts))(*[ACT[a] for a in n.acts])
        keep += [dims, acts]
        nets[k].n_layers = len(n.acts)
        nets[k].dims = dims
        nets[k].acts = acts
        nets[k].theta_offset = int(n.theta_offset)
    terms = (_TermDesc * len(spec.terms))()
    for t, tm in enumerate(spec.terms):
        taps = (_TapDesc * max(1, len(tm.taps)))()
        for i, tp in enumerate(tm.taps):
            taps[i].net, taps[i].out, taps[i].order = int(tp.net), int(tp.out), int(tp.order)
            d = list(tp.dirs) + [0, 0, 0, 0]
            for q in range(4):
                taps[i].dir[q] = int(d[q])
        rows = (C.c_int32 * (len(spec.nets) * MAX_IN))(*([-1] * (len(spec.nets) * MAX_IN)))
        for k, n in enumerate(spec.nets):
            r = tm.net_rows[k] if tm.net_rows is not None and k < len(tm.net_rows) and tm.net_rows[k] is not None \
                else list(range(n.dims[0]))
            for j, v in enumerate(r):
                rows[k * MAX_IN + j] = int(v)
        prog = (_Instr * max(1, len(tm.prog)))()
        for i, ins in enumerate(tm.prog):
            op, a, b, imm = (list(ins) + [0, 0, 0.0])[:4]
            prog[i].op, prog[i].a, prog[i].b, prog[i].imm = OP[op], int(a), int(b), float(imm)
        keep += [taps, rows, prog]
        terms[t].dim = int(tm.dim)
        terms[t].n_taps = len(tm.taps)
        terms[t].taps = taps
        terms[t].net_rows = rows
        terms[t].n_instr = len(tm.prog)
        terms[t].prog = prog
        terms[t].reduction = int(tm.reduction)
        terms[t].scale = float(tm.scale)
    keep += [nets, terms]
    d = _ProblemDesc()
    d.abi_version = ABI_VERSION
    d.dtype = F64 if _np_dtype(spec.dtype) is np.float64 else F32
    d.mode = int(spec.mode)
    d.device = int(spec.device)
    d.n_nets, d.nets = len(spec.nets), nets
    d.n_terms, d.terms = len(spec.terms), terms
    d.n_params, d.param_offset = int(spec.n_params), int(spec.param_offset)
    d.n_theta = int(spec.n_theta)
    return d


class Engine:
    """One engine handle (one GPU rank).  Thin, explicit wrapper over the C ABI."""

    def __init__(self, spec: ProblemSpec):
        self.lib = load_library()
        self.spec = spec
        self.np_dtype = _np_dtype(spec.dtype)
        self.n_terms = len(spec.terms)
        self.n_theta = int(spec.n_theta)
        self._h = C.c_void_p(0)
        desc = build_desc(spec)
        _check(self.lib.pinn_create(C.byref(desc), C.byref(self._h)))
        self._keep_pts = {}

    def close(self):
        if self._h:
            self.lib.pinn_destroy(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- points ---------------------------------------------------------------------------------
    def set_points(self, term: int, dev_pts, n: int, dev_weights=None):
        """Alias device-resident points (torch CUDA tensor or raw pointer), d x n column-major."""
        self._keep_pts[term] = (dev_pts, dev_weights)
        _check(self.lib.pinn_set_points(self._h, term, _ptr(dev_pts), int(n), _ptr(dev_weights)))

    def set_points_host(self, term: int, pts: np.ndarray, weights: Optional[np.ndarray] = None, stream: int = 0):
        """Upload a host (d, n) array (any memory order; converted to d x n column-major)."""
        pts = np.asarray(pts, dtype=self.np_dtype)
        if pts.ndim != 2:
            raise ValueError("points must be a (d, n) matrix")
        if pts.shape[0] != self.spec.terms[term].dim:
            raise ValueError("term %d expects %d rows per point (coordinates + hoisted rows), got %d"
                             % (term, self.spec.terms[term].dim, pts.shape[0]))
        buf = np.asfortranarray(pts)          # column-major: one point = d contiguous scalars
        flat = buf.ravel(order="F")
        w = None if weights is None else np.ascontiguousarray(weights, dtype=self.np_dtype)
        _check(self.lib.pinn_set_points_host(self._h, term, _ptr(flat), int(pts.shape[1]), _ptr(w), C.c_void_p(stream)))

    def set_sampler(self, term: int, n: int, lb, ub, seed: int = 0, stream: int = 0, kind: str = "uniform"):
        """Register a device-side sampler for a term (box lb..ub per point row) and draw the first sample.
        kind "uniform": StochasticTraining; "lhs": Latin hypercube (QuasiRandomTraining's default algorithm)."""
        lb = np.ascontiguousarray(lb, dtype=np.float64); ub = np.ascontiguousarray(ub, dtype=np.float64)
        dim = self.spec.terms[term].dim
        if lb.shape != (dim,) or ub.shape != (dim,):
            raise ValueError("term %d expects %d bounds per side, got %s / %s" % (term, dim, lb.shape, ub.shape))
        _check(self.lib.pinn_set_sampler_ex(self._h, int(term), {"uniform": 0, "lhs": 1}[kind], int(n),
                                            lb.ctypes.data_as(C.POINTER(C.c_double)), ub.ctypes.data_as(C.POINTER(C.c_double)),
                                            C.c_uint64(int(seed) & (2 ** 64 - 1)), C.c_void_p(stream)))
        self._n_pts = getattr(self, "_n_pts", {})
        self._n_pts[int(term)] = int(n)

    def resample(self, stream: int = 0):
        """Draw the next sample of every term that has a device-side sampler."""
        _check(self.lib.pinn_resample(self._h, C.c_void_p(stream)))

    def get_points_host(self, term: int, n: int) -> np.ndarray:
        """Current (d, n) point set of a term, copied from the device."""
        dim = self.spec.terms[term].dim
        buf = np.empty(int(n) * dim, dtype=self.np_dtype)
        _check(self.lib.pinn_get_points_host(self._h, int(term), _ptr(buf)))
        return buf.reshape(int(n), dim).T.copy()

    def set_global_count(self, term: int, n_global: int):
        _check(self.lib.pinn_set_global_count(self._h, term, int(n_global)))

    # -- hot path -------------------------------------------------------------------------------
    def _weights(self, weights):
        if weights is None:
            return None
        w = np.ascontiguousarray(weights, dtype=np.float64)
        if w.shape != (self.n_terms,):
            raise ValueError("need %d term weights" % self.n_terms)
        return w

    def loss_grad_device(self, dev_theta, dev_grad, dev_term_losses, dev_total, weights=None, stream: int = 0):
        w = self._weights(weights)
        wp = w.ctypes.data_as(C.POINTER(C.c_double)) if w is not None else None
        _check(self.lib.pinn_loss_grad(self._h, _ptr(dev_theta), wp, _ptr(dev_grad), _ptr(dev_term_losses),
                                       _ptr(dev_total), C.c_void_p(stream)))

    def loss_grad_host(self, theta: np.ndarray, weights=None, want_grad: bool = True):
        """Host-buffer call: returns (total, term_losses, grad or None)."""
        th = np.ascontiguousarray(theta, dtype=self.np_dtype)
        if th.shape != (self.n_theta,):
            raise ValueError("theta must have length %d" % self.n_theta)
        grad = np.empty(self.n_theta, dtype=self.np_dtype) if want_grad else None
        terms = np.empty(self.n_terms, dtype=self.np_dtype)
        total = np.empty(1, dtype=self.np_dtype)
        w = self._weights(weights)
        wp = w.ctypes.data_as(C.POINTER(C.c_double)) if w is not None else None
        _check(self.lib.pinn_loss_grad_host(self._h, _ptr(th), wp, _ptr(grad), _ptr(terms), _ptr(total)))
        return float(total[0]), terms, grad

    def term_residual_host(self, term: int, theta: np.ndarray, n: int) -> np.ndarray:
        th = np.ascontiguousarray(theta, dtype=self.np_dtype)
        r = np.empty(int(n), dtype=self.np_dtype)
        _check(self.lib.pinn_term_residual_host(self._h, term, _ptr(th), _ptr(r)))
        return r

    def term_grad_stats_host(self, term: int, theta: np.ndarray):
        """(max |g|, mean |g|) of the gradient of term `term`'s unweighted loss (GradientScaleAdaptiveLoss)."""
        th = np.ascontiguousarray(theta, dtype=self.np_dtype)
        mx, mn = C.c_double(0.0), C.c_double(0.0)
        _check(self.lib.pinn_term_grad_stats_host(self._h, int(term), _ptr(th), C.byref(mx), C.byref(mn)))
        return float(mx.value), float(mn.value)

    def term_residual_device(self, term: int, dev_theta, dev_r, stream: int = 0):
        _check(self.lib.pinn_term_residual(self._h, term, _ptr(dev_theta), _ptr(dev_r), C.c_void_p(stream)))

    # -- device-resident Adam loop ----------------------------------------------------------------------
    def adam_begin(self, theta0: np.ndarray, lr=1e-3, beta1=0.9, beta2=0.999, eps=1e-8):
        th = np.ascontiguousarray(theta0, dtype=self.np_dtype)
        _check(self.lib.pinn_adam_begin(self._h, _ptr(th), float(lr), float(beta1), float(beta2), float(eps)))

    def adam_iterate(self, n_steps: int, weights=None):
        """Run n_steps fused iterations on the device; returns (loss, term_losses) of the last evaluated theta."""
        terms = np.empty(self.n_terms, dtype=self.np_dtype)
        total = np.empty(1, dtype=self.np_dtype)
        w = self._weights(weights)
        wp = w.ctypes.data_as(C.POINTER(C.c_double)) if w is not None else None
        _check(self.lib.pinn_adam_iterate(self._h, int(n_steps), wp, _ptr(total), _ptr(terms)))
        return float(total[0]), terms

    def adam_theta(self) -> np.ndarray:
        th = np.empty(self.n_theta, dtype=self.np_dtype)
        _check(self.lib.pinn_adam_theta(self._h, _ptr(th)))
        return th

    # -- multi-GPU --------------------------------------------------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        _check(load_library().pinn_comm_unique_id(C.cast(buf, C.c_void_p)))
        return bytes(buf)

    def comm_init(self, unique_id: bytes, rank: int, nranks: int):
        buf = (C.c_char * 128).from_buffer_copy(unique_id)
        _check(self.lib.pinn_comm_init(self._h, C.cast(buf, C.c_void_p), int(rank), int(nranks)))

    def comm_info(self):
        """(fused_p2p, reason): whether the multi-GPU gradient sum runs inside the fused kernel over peer memory,
        and why not when it fell back to ncclAllReduce."""
        flag = C.c_int32(0)
        why = self.lib.pinn_comm_info(self._h, C.byref(flag))
        return bool(flag.value), (why or b"").decode("utf-8", "replace")

    # -- introspection ------------------------------------------------------------------------------
    def launch_count(self) -> int:
        return int(self.lib.pinn_launch_count(self._h))

    def set_timing(self, on: bool):
        _check(self.lib.pinn_set_timing(self._h, 1 if on else 0))

    def last_kernel_ms(self) -> float:
        return float(self.lib.pinn_last_kernel_ms(self._h))

    def workspace_bytes(self) -> int:
        return int(self.lib.pinn_workspace_bytes(self._h))

    def flops_per_eval(self) -> float:
        return float(self.lib.pinn_flops_per_eval(self._h))
