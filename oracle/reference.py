"""CPU oracle: a float64 restatement of NeuralPDE.jl's PhysicsInformedNN loss path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``oracle/`` is imported by the product package
(``neuralpde.jl_b200``); only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU
baseline legs use it, and only as the checker / the timed CPU baseline.

The real reference is Julia (Lux + Zygote) and cannot run in this image (no ``julia``
binary).  The in-tree arithmetic is restated here line by line:

  Phi                       src/pinn_types.jl:79-90        -> ``phi``
  get_u / numeric_derivative src/pinn_types.jl:442-482      -> ``numeric_derivative``
  get_ε                     src/symbolic_utilities.jl:98-103 -> ``get_eps``
  generated residual        src/symbolic_utilities.jl:132-202,360-370; src/discretize.jl:111-151
                                                            -> ``residual``
  generate_training_sets    src/discretize.jl:185-241       -> ``generate_training_sets``
  get_bounds                src/discretize.jl:299-324       -> ``get_bounds``
  mean(abs2, .)             src/training_strategies.jl:215-221 -> ``term_loss``
  full_loss_function        src/discretize.jl:566-598       -> ``full_loss``
  Zygote gradient           src/discretize.jl:778           -> torch.autograd in float64

The third-party pieces (Lux ``Dense`` = ``act.(W*x .+ b)``, last layer as declared;
ComponentArrays flattening = weight (out x in, column-major) then bias, layer by layer) are
restated from their call sites.  PINNED against the reference's own known-answer tests
(tests/test_oracle_pinning.py): test/Forward/forward__derivatives.jl:25-44 (FD vs exact
gradient / Hessian, atol 1e-8 / 4e-5), test/Forward/forward__ode.jl:44-47 (residual == 2x,
rtol 1e-8), test/Interface/interface__abstract_contracts.jl:57-63 (== 0.5).  The reference
holds NO loss-value or gradient golden vectors for this path, so loss/gradient parity beyond
those fixtures is "parity unpinned" (DESIGN.md section 3).
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np
import sympy as sp
import torch
from sympy.core.function import AppliedUndef

torch.set_default_dtype(torch.float64)

ACTS = {
    "identity": lambda z: z,
    "tanh": torch.tanh,
    "sigmoid": torch.sigmoid,
    "sin": torch.sin,
    "softplus": torch.nn.functional.softplus,
    "swish": lambda z: z * torch.sigmoid(z),
}


# ---- Phi: Lux.Chain of Dense layers on a (d, N) matrix -------------------------------------------
def unpack(theta: torch.Tensor, dims: Sequence[int], offset: int = 0):
    """ComponentArray layout: per layer weight (out x in, column-major) then bias."""
    Ws, bs, o = [], [], offset
    for i in range(len(dims) - 1):
        nin, nout = dims[i], dims[i + 1]
        Ws.append(theta[o:o + nin * nout].reshape(nin, nout).T)   # column-major out x in
        o += nin * nout
        bs.append(theta[o:o + nout])
        o += nout
    return Ws, bs


def phi(x: torch.Tensor, theta: torch.Tensor, dims, acts, offset: int = 0) -> torch.Tensor:
    """``f.smodel(x, θ)``: x is (d, N) features x batch; returns (out, N)."""
    Ws, bs = unpack(theta, dims, offset)
    h = x
    for W, b, a in zip(Ws, bs, acts):
        h = ACTS[a](W @ h + b[:, None])
    return h


# ---- finite-difference derivative (src/pinn_types.jl:445-482) ------------------------------------------
def get_eps(dim: int, der_num: int, dtype, order: int) -> np.ndarray:
    """``get_ε(dim, der_num, T, order)``: eps(T)^(1/(2+order)) on component der_num (0-based here)."""
    e = np.zeros(dim)
    e[der_num] = float(np.finfo(dtype).eps) ** (1.0 / (2 + order))
    return e


def numeric_derivative(u, x: torch.Tensor, eps_list: List[np.ndarray], order: int) -> torch.Tensor:
    """``numeric_derivative(phi, u, x, εs, order, θ)`` with ``u`` closed over (phi, θ)."""
    eps = eps_list[order - 1]
    nz = eps[eps != 0.0]
    inv = 1.0 / nz[0]
    e = torch.tensor(eps, dtype=x.dtype).reshape(-1, 1)
    if order > 4 or any(not np.array_equal(v, eps_list[0]) for v in eps_list[:order]):
        return (numeric_derivative(u, x + e, eps_list[:order - 1], order - 1)
                - numeric_derivative(u, x - e, eps_list[:order - 1], order - 1)) * inv / 2
    if order == 4:
        return (u(x + 2 * e) - 4 * u(x + e) + 6 * u(x) - 4 * u(x - e) + u(x - 2 * e)) * inv ** 4
    if order == 3:
        return (u(x + 2 * e) - 2 * u(x + e) + 2 * u(x - e) - u(x - 2 * e)) * inv ** 3 / 2
    if order == 2:
        return (u(x + e) + u(x - e) - 2 * u(x)) * inv ** 2
    if order == 1:
        return (u(x + e) - u(x - e)) * inv / 2
    raise RuntimeError("This shouldn't happen!")


# ---- exact taps (closed-form forward-mode propagation) ---------------------------------
def _act_derivs(name: str, z: torch.Tensor):
    if name == "identity":
        return z, torch.ones_like(z), torch.zeros_like(z)
    if name == "tanh":
        t = torch.tanh(z)
        s = 1 - t * t
        return t, s, -2 * t * s
    if name == "sigmoid":
        g = torch.sigmoid(z)
        g1 = g * (1 - g)
        return g, g1, g1 * (1 - 2 * g)
    if name == "sin":
        return torch.sin(z), torch.cos(z), -torch.sin(z)
    if name == "softplus":
        g = torch.sigmoid(z)
        return torch.nn.functional.softplus(z), g, g * (1 - g)
    if name == "swish":
        g = torch.sigmoid(z)
        g1 = g * (1 - g)
        g2 = g1 * (1 - 2 * g)
        return z * g, g + z * g1, 2 * g1 + z * g2
    raise ValueError(name)


def exact_tap(x: torch.Tensor, theta: torch.Tensor, dims, acts, offset: int, dirs: Tuple[int, ...]) -> torch.Tensor:
    """Value (dirs=()), first (dirs=(d,)) or second (dirs=(d, e)) partial derivative of the
    network output, propagated in closed form through every layer; orders above 2 by nested autograd."""
    if len(dirs) > 2:
        # higher orders: nested reverse-mode differentiation of the network with respect to its input (independent of the
        # closed-form propagation the engine uses); columns are independent, so summing over the batch is exact
        xg = x.detach().clone().requires_grad_(True)
        v = phi(xg, theta, dims, acts, offset)
        for d_ in dirs:
            (g,) = torch.autograd.grad(v.sum(), xg, create_graph=True)
            v = g[d_:d_ + 1, :]
        return v
    Ws, bs = unpack(theta, dims, offset)
    n = x.shape[1]
    h = x
    hd = []
    for d_ in dirs:
        e = torch.zeros(dims[0], n, dtype=x.dtype)
        e[d_, :] = 1.0
        hd.append(e)
    hdd = torch.zeros(dims[0], n, dtype=x.dtype) if len(dirs) == 2 else None
    for W, b, a in zip(Ws, bs, acts):
        z = W @ h + b[:, None]
        zd = [W @ v for v in hd]
        val, d1, d2 = _act_derivs(a, z)
        if hdd is not None:
            zdd = W @ hdd
            hdd = d1 * zdd + d2 * zd[0] * zd[1]
        hd = [d1 * v for v in zd]
        h = val
    if len(dirs) == 0:
        return h
    if len(dirs) == 1:
        return hd[0]
    return hdd


# ---- variable bookkeeping (src/symbolic_utilities.jl:401-526), restated independently ----------------------
def _depvar_names(dvs) -> List[str]:
    return [d.func.__name__ if isinstance(d, AppliedUndef) else str(d) for d in dvs]


def _depvar_inputs(ivs, dvs) -> Dict[str, List[str]]:
    return {(d.func.__name__ if isinstance(d, AppliedUndef) else str(d)):
            ([str(a) for a in d.args] if isinstance(d, AppliedUndef) else [str(v) for v in ivs]) for d in dvs}


def _first_apps(eq, names: List[str]) -> Dict[str, AppliedUndef]:
    first: Dict[str, AppliedUndef] = {}
    for side in (eq.lhs, eq.rhs):
        for node in sp.preorder_traversal(side):
            if isinstance(node, AppliedUndef) and node.func.__name__ in names:
                first.setdefault(node.func.__name__, node)
    return first


def get_argument(eqs, ivs, dvs) -> List[list]:
    names = _depvar_names(dvs)
    out = []
    for eq in eqs:
        first = _first_apps(eq, names)
        args, seen = [], set()
        for nm in names:
            if nm in first:
                for a in first[nm].args:
                    if isinstance(a, sp.Symbol):
                        if str(a) not in seen:
                            seen.add(str(a))
                            args.append(str(a))
                    else:
                        args.append(float(a))
        out.append(args)
    return out


def eq_indvars(eq, ivs, dvs) -> List[str]:
    names = _depvar_names(dvs)
    ins = _depvar_inputs(ivs, dvs)
    first = _first_apps(eq, names)
    out: List[str] = []
    for nm in names:
        if nm in first:
            for v in ins[nm]:
                if v not in out:
                    out.append(v)
    return out


# ---- training sets (src/discretize.jl:185-241, :299-324) ----------------------------------------------------
def _range(lo, dx, hi):
    n = int(np.floor((hi - lo) / dx + 1e-10)) + 1
    return lo + dx * np.arange(n)


def _product(spans):
    grids = np.meshgrid(*spans, indexing="ij")
    return np.stack([g.ravel(order="F") for g in grids], axis=0)


def generate_training_sets(domains, dx, eqs, bcs, ivs, dvs):
    """Grid sets.  `dif` (src/discretize.jl:214-218) is filled from ``get_variables`` -- symbols
    only -- so it is always empty and ``setdiff`` removes nothing: the PDE set is the full grid."""
    dxs = list(dx) if isinstance(dx, (list, tuple, np.ndarray)) else [dx] * len(domains)
    span = {str(d.variables): _range(d.domain.lo, h, d.domain.hi) for d, h in zip(domains, dxs)}

    def build(eq_list):
        return [_product([span[a] if isinstance(a, str) else np.array([a]) for a in args])
                for args in get_argument(eq_list, ivs, dvs)]

    return build(eqs), build(bcs)


def get_bounds(domains, eqs, bcs, ivs, dvs, points: int):
    dx = 1.0 / points
    span = {str(d.variables): (d.domain.lo + dx, d.domain.hi - dx) for d in domains}

    def build(eq_list):
        out = []
        for args in get_argument(eq_list, ivs, dvs):
            lo = np.array([span[a][0] if isinstance(a, str) else a for a in args], dtype=np.float64)
            hi = np.array([span[a][1] if isinstance(a, str) else a for a in args], dtype=np.float64)
            out.append((lo, hi))
        return out

    return build(eqs), build(bcs)


# ---- generated residual -----------------------------------------------------------------------------------
class Problem:
    """Everything the generated loss closures close over."""

    def __init__(self, pde_system, chains: Sequence[Tuple[Sequence[int], Sequence[str]]], param_estim: bool = False,
                 eltype=np.float64, derivative: str = "fd"):
        self.sys = pde_system
        self.names = _depvar_names(pde_system.dvs)
        self.inputs = _depvar_inputs(pde_system.ivs, pde_system.dvs)
        self.chains = list(chains)                       # (dims, acts) per depvar
        self.offsets, o = [], 0
        for dims, _ in self.chains:
            self.offsets.append(o)
            o += sum(dims[i] * dims[i + 1] + dims[i + 1] for i in range(len(dims) - 1))
        self.n_net = o
        self.param_names = [str(p) for p in pde_system.ps]
        self.param_estim = param_estim
        self.defaults = {str(k): float(v) for k, v in pde_system.defaults.items()}
        self.n_theta = o + (len(self.param_names) if param_estim else 0)
        self.eltype = eltype                             # sets ε = eps(eltype)^(1/(2+order))
        self.derivative = derivative                     # "fd" (reference) | "exact"

    # u(cord_k, θ_k, phi_k) and derivative(phi_k, u, cord_k, εs, order, θ_k)
    def _u(self, k: int, theta):
        dims, acts = self.chains[k]
        return lambda c: phi(c, theta, dims, acts, self.offsets[k])

    def _eval(self, e, env: Dict[str, torch.Tensor], cords: Dict[str, torch.Tensor], theta) -> torch.Tensor:
        if isinstance(e, sp.Symbol):
            nm = str(e)
            if nm in env:
                return env[nm]
            if nm in self.param_names:
                i = self.param_names.index(nm)
                if self.param_estim:
                    return theta[self.n_net + i:self.n_net + i + 1].reshape(1, 1)    # θ.p[i:i]
                return torch.tensor(self.defaults[nm]).reshape(1, 1)
            raise KeyError(nm)
        if isinstance(e, AppliedUndef):
            k = self.names.index(e.func.__name__)
            return self._u(k, theta)(cords[e.func.__name__])
        if isinstance(e, sp.Subs):
            return self._eval(e.args[0], env, cords, theta)
        if isinstance(e, sp.Derivative):
            dv, inner = [], e
            while isinstance(inner, sp.Derivative):
                for v, n in inner.variable_count:
                    dv += [str(v)] * int(n)
                inner = inner.expr
            if isinstance(inner, sp.Subs):
                inner = inner.args[0]
                while isinstance(inner, sp.Derivative):
                    for v, n in inner.variable_count:
                        dv += [str(v)] * int(n)
                    inner = inner.expr
            nm = inner.func.__name__
            k = self.names.index(nm)
            slots = self.inputs[nm]
            order = len(dv)
            undv = [slots.index(v) for v in dv]
            if self.derivative == "exact":
                dims, acts = self.chains[k]
                return exact_tap(cords[nm], theta, dims, acts, self.offsets[k], tuple(undv))
            eps = [get_eps(len(slots), d_, self.eltype, order) for d_ in range(len(slots))]
            return numeric_derivative(self._u(k, theta), cords[nm], [eps[d_] for d_ in undv], order)
        if isinstance(e, (sp.Number, sp.NumberSymbol)) or (e.is_number and not e.free_symbols):
            return torch.tensor(float(e)).reshape(1, 1)
        if isinstance(e, sp.Add):
            out = self._eval(e.args[0], env, cords, theta)
            for a in e.args[1:]:
                out = out + self._eval(a, env, cords, theta)
            return out
        if isinstance(e, sp.Mul):
            out = self._eval(e.args[0], env, cords, theta)
            for a in e.args[1:]:
                out = out * self._eval(a, env, cords, theta)
            return out
        if isinstance(e, sp.Pow):
            b = self._eval(e.args[0], env, cords, theta)
            if e.args[1].is_Integer:
                return b ** int(e.args[1])
            return b ** self._eval(e.args[1], env, cords, theta)
        fn = {sp.sin: torch.sin, sp.cos: torch.cos, sp.exp: torch.exp, sp.log: torch.log, sp.tanh: torch.tanh,
              sp.Abs: torch.abs, sp.tan: torch.tan, sp.cosh: torch.cosh, sp.sinh: torch.sinh}
        for f, t in fn.items():
            if isinstance(e, f):
                return t(self._eval(e.args[0], env, cords, theta))
        raise NotImplementedError("oracle: %s" % type(e).__name__)

    @staticmethod
    def _expand(ex):
        ex = sp.sympify(ex)
        d = ex.doit()
        return ex if d == 0 else d            # the _iszero fallback of parse_equation

    def residual(self, eq, cord: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
        """The generated ``(cord, θ) -> lhs .- rhs`` closure (src/discretize.jl:111-151)."""
        ivs = eq_indvars(eq, self.sys.ivs, self.sys.dvs)
        env = {v: cord[i:i + 1, :] for i, v in enumerate(ivs)}           # x = cord[[i], :]
        cords = {}
        first = _first_apps(eq, self.names)
        for nm in first:
            cords[nm] = torch.cat([env[v] for v in self.inputs[nm]], dim=0)   # cord_k = vcat(...)
        lhs = self._eval(self._expand(eq.lhs), env, cords, theta)
        rhs = self._eval(self._expand(eq.rhs), env, cords, theta)
        r = lhs - rhs
        return r.expand(1, cord.shape[1]) if r.shape[1] == 1 else r

    def term_loss(self, eq, cord, theta, weights: Optional[torch.Tensor] = None, scale: float = 1.0):
        r = self.residual(eq, cord, theta)
        if weights is None:
            return torch.mean(r * r)                                    # mean(abs2, .)
        return scale * torch.sum(weights * r[0] ** 2)

    def full_loss(self, theta, pde_sets, bc_sets, pde_w=None, bc_w=None, extra=None, qweights=None, qscales=None):
        """``Σ w_pde L_pde + Σ w_bc L_bc (+ w_add * additional)`` (src/discretize.jl:582-598).
        Returns (total, [term losses])."""
        eqs = list(self.sys.eqs) + list(self.sys.bcs)
        sets = list(pde_sets) + list(bc_sets)
        w = ([1.0] * len(self.sys.eqs) if pde_w is None else list(pde_w)) + \
            ([1.0] * len(self.sys.bcs) if bc_w is None else list(bc_w))
        terms = []
        for i, (eq, s) in enumerate(zip(eqs, sets)):
            qw = None if qweights is None or qweights[i] is None else torch.as_tensor(qweights[i])
            sc = 1.0 if qscales is None else qscales[i]
            terms.append(self.term_loss(eq, torch.as_tensor(s, dtype=torch.float64), theta, qw, sc))
        total = sum(wi * t for wi, t in zip(w, terms))
        if extra is not None:
            w_add, fn = extra
            t = fn(theta)
            terms.append(t)
            total = total + w_add * t
        return total, terms

    def loss_and_grad(self, theta_np: np.ndarray, pde_sets, bc_sets, **kw):
        theta = torch.tensor(np.asarray(theta_np, dtype=np.float64), requires_grad=True)
        total, terms = self.full_loss(theta, pde_sets, bc_sets, **kw)
        (g,) = torch.autograd.grad(total, theta)
        return float(total.detach()), np.array([float(t.detach()) for t in terms]), g.numpy().copy()

    def data_loss(self, depvar: str, X: np.ndarray, y: np.ndarray):
        """``mean(abs2, u_k(X) .- y)`` as an additional_loss closure."""
        k = self.names.index(depvar)
        Xt, yt = torch.as_tensor(X, dtype=torch.float64), torch.as_tensor(y, dtype=torch.float64).reshape(1, -1)
        return lambda theta: torch.mean((self._u(k, theta)(Xt) - yt) ** 2)
