"""Lowering of one equation / boundary condition to the engine's residual IR.

Input: the equation after ``expand_derivatives`` (so every derivative chain ends directly on
a dependent-variable call -- the invariant the reference's tap extractor relies on,
src/symbolic_utilities.jl:160-174).  Output: the tap list (pure partial derivatives of one
network each) and an SSA program computing ``lhs - rhs`` per point from taps, coordinate
rows and parameters -- the same quantity the reference's generated function returns
(src/symbolic_utilities.jl:360-370, src/discretize.jl:126-151).

In deployment this is what the Julia shim does with ``pinnrep.symbolic_*_loss_functions``
(INTEGRATION.md); it lives here in Python because Julia is absent from the build image.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Optional, Tuple

import sympy as sp
from sympy.core.function import AppliedUndef

from .engine import TapSpec, TermSpec, REDUCE_MEAN
from .symbolic import Equation, VarInfo, eq_indvars, expand_derivatives


class LoweringError(ValueError):
    pass


@dataclass
class LoweredTerm:
    taps: List[TapSpec]
    prog: List[tuple]
    indvars: List[str]               # row i of the term's point matrix is this variable
    net_rows: List[Optional[List[int]]]
    # coordinate-only subexpressions hoisted out of the per-step program: evaluated once per point set
    # (float64, on the host) and appended as extra rows dim, dim+1, ... of the point matrix
    extra_exprs: List[sp.Expr] = None

    @property
    def dim(self) -> int:
        return len(self.indvars) + len(self.extra_exprs or [])

    def augment(self, pts):
        """(d, N) coordinates -> (d + n_extra, N) with the hoisted rows appended."""
        import numpy as np
        pts = np.asarray(pts)
        if not self.extra_exprs:
            return pts
        if pts.shape[0] != len(self.indvars):
            raise ValueError("expected %d coordinate rows, got %d" % (len(self.indvars), pts.shape[0]))
        rows64 = [pts[i].astype(np.float64) for i in range(pts.shape[0])]
        fns = getattr(self, "_extra_fns", None)
        if fns is None:          # compiled once per term: resampling strategies call augment on every loss evaluation
            syms = [sp.Symbol(v, real=True) for v in self.indvars]
            fns = [sp.lambdify(syms, e, "numpy") for e in self.extra_exprs]
            self._extra_fns = fns
        extra = []
        for f in fns:
            extra.append(np.broadcast_to(np.asarray(f(*rows64), dtype=np.float64), rows64[0].shape))
        return np.concatenate([pts, np.stack(extra).astype(pts.dtype)], axis=0)


MAX_DIM = 8   # PINN_MAX_DIM


def _hoist_coordinate_terms(expr, coord_names, blocked_names, extras, base_dim):
    """Replace maximal coordinate-only subexpressions (no network tap, no trainable parameter) by fresh
    symbols __c<i>; trivial ones (a bare symbol / number, or fewer than 2 operations) stay inline."""
    def has_net(e):
        return e.has(AppliedUndef) or e.has(sp.Derivative) or any(str(q) in blocked_names for q in e.free_symbols)

    def rec(e):
        if isinstance(e, (AppliedUndef, sp.Derivative, sp.Subs)):
            return e
        if not has_net(e):
            names = {str(q) for q in e.free_symbols}
            if names and names <= coord_names and not isinstance(e, sp.Symbol) and e.count_ops() >= 2 \
                    and base_dim + len(extras) < MAX_DIM:
                for i, old in enumerate(extras):
                    if old == e:
                        return sp.Symbol("__c%d" % i, real=True)
                extras.append(e)
                return sp.Symbol("__c%d" % (len(extras) - 1), real=True)
            return e
        if isinstance(e, (sp.Add, sp.Mul)):
            pure = [a for a in e.args if not has_net(a)]
            rest = [rec(a) for a in e.args if has_net(a)]
            if pure:
                rest.append(rec(e.func(*pure)))
            return e.func(*rest, evaluate=False) if len(rest) > 1 else rest[0]
        if e.args:
            return e.func(*[rec(a) for a in e.args])
        return e

    return rec(sp.sympify(expr))


class _Emitter:
    def __init__(self, vi: VarInfo, rows: List[str], param_index: Dict[str, int], param_values: Dict[str, float]):
        self.vi = vi
        self.rows = rows
        self.param_index = param_index
        self.param_values = param_values
        self.prog: List[tuple] = []
        self.taps: List[TapSpec] = []
        self._tap_ids: Dict[tuple, int] = {}
        self._cse: Dict[object, int] = {}

    def _push(self, op, a=0, b=0, imm=0.0) -> int:
        key = (op, a, b, float(imm))
        if key in self._cse:
            return self._cse[key]
        self.prog.append((op, int(a), int(b), float(imm)))
        self._cse[key] = len(self.prog) - 1
        return len(self.prog) - 1

    def const(self, v: float) -> int:
        return self._push("const", imm=float(v))

    def tap(self, depvar: str, dirs: Tuple[int, ...]) -> int:
        net = self.vi.dict_depvars[depvar]
        dirs = tuple(sorted(dirs))
        key = (net, dirs)
        if key not in self._tap_ids:
            if len(dirs) > 3 or (len(dirs) == 3 and len(set(dirs)) != 1):
                raise LoweringError(
                    "derivative of order %d of %s along %s: this engine propagates exact taps up to order 2 in any "
                    "directions and pure third derivatives (the reference's order-4 stencil and the recursive mixed "
                    "forms, src/pinn_types.jl:454-474, are not covered)" % (len(dirs), depvar, list(dirs)))
            self._tap_ids[key] = len(self.taps)
            self.taps.append(TapSpec(net=net, order=len(dirs), dirs=dirs))
        return self._push("tap", a=self._tap_ids[key])

    # -- expression walk -----------------------------------------------------------------------
    def emit(self, e: sp.Expr) -> int:
        e = sp.sympify(e)
        if isinstance(e, (sp.Number, sp.NumberSymbol)) or e.is_number and not e.free_symbols and not e.has(AppliedUndef):
            return self.const(float(e))
        if isinstance(e, sp.Symbol):
            name = str(e)
            if name in self.rows:
                return self._push("coord", a=self.rows.index(name))
            if name.startswith("__c"):
                return self._push("coord", a=len(self.rows) + int(name[3:]))
            if name in self.param_index:
                return self._push("param", a=self.param_index[name])
            if name in self.param_values:
                return self.const(self.param_values[name])
            if name in self.vi.dict_indvars:
                raise LoweringError(
                    "independent variable %s is not an input of any dependent variable in this equation" % name)
            raise LoweringError("unknown symbol %s (not an independent variable or a parameter with a default)" % name)
        if isinstance(e, AppliedUndef):
            name = e.func.__name__
            if name not in self.vi.dict_depvars:
                raise LoweringError("unknown function %s" % name)
            if len(e.args) != len(self.vi.dict_depvar_input[name]):
                raise LoweringError("%s called with %d arguments, declared with %d"
                                    % (name, len(e.args), len(self.vi.dict_depvar_input[name])))
            return self.tap(name, ())
        if isinstance(e, sp.Subs):
            # Subs(Derivative(u(x,y), x), x, 0): the evaluation point lives in the data
            return self.emit(e.args[0])
        if isinstance(e, sp.Derivative):
            dvars: List[str] = []
            inner = e
            while isinstance(inner, sp.Derivative):
                for v, n in inner.variable_count:
                    dvars += [str(v)] * int(n)
                inner = inner.expr
            if isinstance(inner, sp.Subs):
                inner = inner.args[0]
                while isinstance(inner, sp.Derivative):
                    for v, n in inner.variable_count:
                        dvars += [str(v)] * int(n)
                    inner = inner.expr
            if not (isinstance(inner, AppliedUndef) and inner.func.__name__ in self.vi.dict_depvars):
                raise LoweringError("derivative of a non-network expression survived expand_derivatives: %s" % e)
            name = inner.func.__name__
            slots = self.vi.dict_depvar_input[name]
            dirs = []
            for v in dvars:
                if v not in slots:
                    raise LoweringError("derivative of %s with respect to %s, which is not one of its inputs %s"
                                        % (name, v, slots))
                dirs.append(slots.index(v))
            return self.tap(name, tuple(dirs))
        if isinstance(e, sp.Add):
            terms = list(e.args)
            acc = None
            for t in terms:
                coeff, rest = t.as_coeff_Mul()
                if coeff == -1 and rest != 1 and acc is not None:
                    acc = self._push("sub", a=acc, b=self.emit(rest))
                else:
                    v = self.emit(t)
                    acc = v if acc is None else self._push("add", a=acc, b=v)
            return acc
        if isinstance(e, sp.Mul):
            coeff, rest = e.as_coeff_Mul()
            if coeff == -1 and rest != 1:
                return self._push("neg", a=self.emit(rest))
            num, den = [], []
            for f in e.args:
                if isinstance(f, sp.Pow) and f.exp.is_number and f.exp.is_negative:
                    den.append(sp.Pow(f.base, -f.exp))
                else:
                    num.append(f)
            acc = None
            for f in num:
                v = self.emit(f)
                acc = v if acc is None else self._push("mul", a=acc, b=v)
            if acc is None:
                acc = self.const(1.0)
            for f in den:
                acc = self._push("div", a=acc, b=self.emit(f))
            return acc
        if isinstance(e, sp.Pow):
            base, ex = e.args
            if base == sp.E:
                return self._push("exp", a=self.emit(ex))
            if ex.is_Integer:
                n = int(ex)
                if n == 2:
                    b = self.emit(base)
                    return self._push("mul", a=b, b=b)
                return self._push("powi", a=self.emit(base), imm=float(n))
            if ex == sp.Rational(1, 2):
                return self._push("sqrt", a=self.emit(base))
            if ex == sp.Rational(-1, 2):
                return self._push("div", a=self.const(1.0), b=self._push("sqrt", a=self.emit(base)))
            return self._push("pow", a=self.emit(base), b=self.emit(ex))
        unary = {sp.sin: "sin", sp.cos: "cos", sp.exp: "exp", sp.log: "log", sp.tanh: "tanh", sp.Abs: "abs"}
        for f, op in unary.items():
            if isinstance(e, f):
                return self._push(op, a=self.emit(e.args[0]))
        if isinstance(e, sp.tan):
            a = self.emit(e.args[0])
            return self._push("div", a=self._push("sin", a=a), b=self._push("cos", a=a))
        if isinstance(e, sp.cosh) or isinstance(e, sp.sinh):
            a = self.emit(e.args[0])
            ep = self._push("exp", a=a)
            en = self._push("exp", a=self._push("neg", a=a))
            s = self._push("add" if isinstance(e, sp.cosh) else "sub", a=ep, b=en)
            return self._push("mul", a=s, b=self.const(0.5))
        raise LoweringError("unsupported expression node %s in %s" % (type(e).__name__, e))


def lower_equation(eq: Equation, vi: VarInfo, param_index: Optional[Dict[str, int]] = None,
                   param_values: Optional[Dict[str, float]] = None, hoist: bool = False) -> LoweredTerm:
    """Equation -> taps + residual program (``lhs - rhs``)."""
    rows = eq_indvars(eq, vi)
    em = _Emitter(vi, rows, param_index or {}, param_values or {})
    lhs = expand_derivatives(eq.lhs)
    rhs = expand_derivatives(eq.rhs)
    extras: List[sp.Expr] = []
    if hoist:
        # trainable parameters block hoisting; parameters with fixed defaults are substituted first
        subs = {sp.Symbol(k, real=True): v for k, v in (param_values or {}).items()}
        lhs = _hoist_coordinate_terms(lhs.subs(subs), set(rows), set(param_index or {}), extras, len(rows))
        rhs = _hoist_coordinate_terms(rhs.subs(subs), set(rows), set(param_index or {}), extras, len(rows))
    a = em.emit(lhs)
    b = em.emit(rhs)
    em.prog.append(("sub", a, b, 0.0))      # not CSE'd: must be the last instruction
    if not em.taps:
        raise LoweringError("equation %s contains no dependent variable: nothing to train on" % (eq,))
    net_rows: List[Optional[List[int]]] = []
    for name in vi.depvars:
        ins = vi.dict_depvar_input[name]
        if all(v in rows for v in ins):
            net_rows.append([rows.index(v) for v in ins])
        else:
            net_rows.append(None)
    return LoweredTerm(em.taps, em.prog, rows, net_rows, extras)


def term_spec(lt: LoweredTerm, reduction: int = REDUCE_MEAN, scale: float = 1.0) -> TermSpec:
    return TermSpec(dim=lt.dim, taps=lt.taps, prog=lt.prog, net_rows=lt.net_rows,
                    reduction=reduction, scale=scale)
