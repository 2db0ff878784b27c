"""Shared helpers for the parity tests: run a Config through the oracle and the engine."""
import numpy as np

import neuralpde_jl_b200 as npde
from oracle import reference as R


def oracle_eval(cfg, theta64, derivative="exact", point_sets=None, quad=None):
    """Float64 oracle loss / term losses / gradient of a Config at the given point sets
    (default: the Grid sets the oracle itself generates)."""
    sys_ = cfg.pde_system
    prob = R.Problem(sys_, cfg.chain_specs(), param_estim=cfg.param_estim, derivative=derivative)
    n_pde = len(sys_.eqs)
    if point_sets is None:
        ps, bs = R.generate_training_sets(sys_.domain, cfg.strategy.dx, sys_.eqs, sys_.bcs, sys_.ivs, sys_.dvs)
    else:
        ps, bs = point_sets[:n_pde], point_sets[n_pde:n_pde + len(sys_.bcs)]
    kw = {}
    if quad is not None:
        kw["qweights"], kw["qscales"] = quad
    if cfg.additional_loss is not None:
        a = cfg.additional_loss
        kw["extra"] = (1.0, prob.data_loss(a.depvar, a.points, a.values))
    return prob.loss_and_grad(theta64, ps, bs, **kw)


def engine_eval(cfg, dtype, mode="ffma", want_grad=True):
    disc = cfg.discretization(dtype=dtype, mode=mode)
    rep = npde.symbolic_discretize(cfg.pde_system, disc)
    theta = rep.flat_init_params
    total, terms, grad = rep.engine.loss_grad_host(theta, None, want_grad)
    return rep, total, terms, grad


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


def load_golden(name):
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    n = int(g["n_sets"])
    sets = [g["set_%d" % i] for i in range(n)]
    qw = [g["qw_%d" % i] for i in range(n)] if "qw_0" in g else None
    return g, sets, qw


def engine_eval_sets(cfg, dtype, sets, qw=None, mode="ffma", want_grad=True, theta=None):
    """Engine loss / terms / grad of a Config at explicit point sets (overrides the strategy's sets)."""
    disc = cfg.discretization(dtype=dtype, mode=mode)
    rep = npde.symbolic_discretize(cfg.pde_system, disc)
    th = rep.flat_init_params if theta is None else np.asarray(theta, dtype=dtype)
    for i, s in enumerate(sets):
        rep.set_points(i, s, None if qw is None else qw[i])
    total, terms, grad = rep.engine.loss_grad_host(th, None, want_grad)
    return rep, total, terms, grad
