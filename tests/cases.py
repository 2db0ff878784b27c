"""Parity cases shared by the golden generator, the CPU tests and the GPU tests."""
import numpy as np
import sympy as sp

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from neuralpde_jl_b200.configs import Config, mlp
from neuralpde_jl_b200.pinn import Chain, Dense
from neuralpde_jl_b200.strategies import (GridTraining, QuadratureTraining, QuasiRandomTraining, StochasticTraining,
                                          gauss_legendre_box, generate_quasi_random_points, generate_random_points,
                                          generate_training_sets, get_bounds)
from neuralpde_jl_b200.symbolic import Differential, Eq, In, PDESystem, get_vars, parameters, variables


def mixed_derivative_case() -> Config:
    """u_xx + u_xy - 2 u_yy = -1 style equation (reference test/NNPDE1/nnpde__pde_vi_pde_with_mixed_derivative.jl)."""
    x, y = parameters("x y")
    u = variables("u")
    Dx, Dy = Differential(x), Differential(y)
    Dxx, Dyy = Dx ** 2, Dy ** 2
    eq = Eq(Dxx(u(x, y)) + Dx(Dy(u(x, y))) - 2 * Dyy(u(x, y)), -1.0)
    bcs = [Eq(u(x, 0), x), Eq(Dy(u(x, 0)), x), Eq(u(0, y), sp.exp(y) - 1), Eq(u(1, y), sp.exp(y))]
    sys_ = PDESystem(eq, bcs, [In(x, 0.0, 1.0), In(y, 0.0, 1.0)], [x, y], [u(x, y)])
    return Config("mixed_derivative", sys_, [mlp(2, 16, 2, "sigmoid")], GridTraining(0.1), n_pde_points=121)


def neumann_sin_case() -> Config:
    """1-D problem with a Neumann bc and sin / softplus activations (value + first-derivative channels)."""
    x = parameters("x")
    u = variables("u")
    Dx = Differential(x)
    eq = Eq((Dx ** 2)(u(x)) + u(x) * Dx(u(x)), sp.cos(2 * x) * sp.exp(-x) + x ** 3 / (1 + x ** 2))
    bcs = [Eq(u(0.0), 0.5), Eq(Dx(u(1.0)), -0.25)]
    sys_ = PDESystem(eq, bcs, [In(x, 0.0, 1.0)], [x], [u(x)])
    chain = Chain(Dense(1, 16, "sin"), Dense(16, 16, "softplus"), Dense(16, 16, "swish"), Dense(16, 1))
    return Config("neumann_sin", sys_, [chain], GridTraining(1.0 / 63), n_pde_points=64)


def poisson1d_wide_case() -> Config:
    """1-D Poisson with a Neumann end on a 1 -> 128 -> 64 -> 128 -> 1 network: the 128-wide tensor path with mixed
    64 / 128 layer widths and 3 / 2 / 1 propagated channels (PDE, Neumann, Dirichlet terms)."""
    x = parameters("x")
    u = variables("u")
    Dx = Differential(x)
    eq = Eq((Dx ** 2)(u(x)), -sp.pi ** 2 * sp.sin(sp.pi * x))
    bcs = [Eq(u(0.0), 0.0), Eq(Dx(u(1.0)), -sp.pi)]
    sys_ = PDESystem(eq, bcs, [In(x, 0.0, 1.0)], [x], [u(x)])
    chain = Chain(Dense(1, 128, "tanh"), Dense(128, 64, "tanh"), Dense(64, 128, "tanh"), Dense(128, 1))
    return Config("poisson1d_wide", sys_, [chain], GridTraining(1.0 / 299), n_pde_points=300)


def third_order_ode_case() -> Config:
    """1-D third-order ODE u_xxx + u u_x = cos(pi x) (the equation of reference test/NNPDE1/nnpde__pde_iii_3rd_order_ode.jl
    written with Dxxx = Differential(x)^3 directly; stencil src/pinn_types.jl:469-474) with Dirichlet, Neumann and
    second-derivative boundary terms; tanh / sigmoid / swish layers exercise the fourth activation derivative of the
    reverse sweep."""
    x = parameters("x")
    u = variables("u")
    Dx = Differential(x)
    eq = Eq((Dx ** 3)(u(x)) + u(x) * Dx(u(x)), sp.cos(sp.pi * x))
    bcs = [Eq(u(0.0), 0.0), Eq(u(1.0), -1.0), Eq(Dx(u(1.0)), 1.0), Eq((Dx ** 2)(u(0.0)), 0.25)]
    sys_ = PDESystem(eq, bcs, [In(x, 0.0, 1.0)], [x], [u(x)])
    chain = Chain(Dense(1, 16, "tanh"), Dense(16, 16, "sigmoid"), Dense(16, 16, "swish"), Dense(16, 1))
    return Config("third_order_ode", sys_, [chain], GridTraining(1.0 / 49), n_pde_points=50)


def third_order_2d_case() -> Config:
    """u_yyy + u_xx + u_x u_y = sin(x) y on [0,1]^2: a pure third derivative along the SECOND input next to a second
    derivative along the first (channel ordering), sin / softplus layers."""
    x, y = parameters("x y")
    u = variables("u")
    Dx, Dy = Differential(x), Differential(y)
    eq = Eq((Dy ** 3)(u(x, y)) + (Dx ** 2)(u(x, y)) + Dx(u(x, y)) * Dy(u(x, y)), sp.sin(x) * y)
    bcs = [Eq(u(0, y), y), Eq((Dy ** 2)(u(x, 0)), x), Eq(Dy(u(x, 1)), 0.5)]
    sys_ = PDESystem(eq, bcs, [In(x, 0.0, 1.0), In(y, 0.0, 1.0)], [x, y], [u(x, y)])
    chain = Chain(Dense(2, 16, "sin"), Dense(16, 16, "softplus"), Dense(16, 1))
    return Config("third_order_2d", sys_, [chain], GridTraining(0.1), n_pde_points=121)


CASES = {
    "cfg1": lambda: configs.config1(),
    "cfg2_small": lambda: configs.config2(n=24, width=16, hidden=2),
    "cfg3_small": lambda: configs.config3(points=512, bcs_points=96, width=32, hidden=3),
    "cfg4_tiny": lambda: configs.config4(nodes=4, bc_nodes=3, width=16, hidden=2),
    "cfg5_small": lambda: configs.config5(points=384, bcs_points=64, n_obs=80, width=16, hidden=2),
    "burgers_wide": lambda: configs.config3(points=700, bcs_points=150, width=128, hidden=3),
    "poisson1d_wide": poisson1d_wide_case,
    "cfg5_wide": lambda: configs.config5(points=500, bcs_points=70, n_obs=90, width=128, hidden=3),
    "mixed": mixed_derivative_case,
    "neumann_sin": neumann_sin_case,
    "third_order_ode": third_order_ode_case,
    "third_order_2d": third_order_2d_case,
}


# the shapes BASELINE.json names (goldens: tests/golden/make_golden_full.py; theta / points regenerated from seeds)
FULL_CASES = {
    "cfg2_full": lambda: configs.config2(),                                                        # 128^2, 4x64
    "cfg3_full": lambda: configs.config3(),                                                        # 65 536 + 3 x 4 096, 5x128
    "cfg5_full": lambda: configs.config5(points=65536, bcs_points=4096, n_obs=1024),               # 4x128, 6 channels, data + theta.p
    "cfg4_w256": lambda: configs.config4(nodes=6, bc_nodes=4, width=256, hidden=2),                # 4 nets, 256-wide, 7 channels
}


def point_sets(cfg: Config, seed: int = 11):
    """Deterministic (d, N) float64 point sets [pde..., bc...] (+ quadrature weights / scales) of a case."""
    sys_ = cfg.pde_system
    vi = get_vars(sys_.ivs, sys_.dvs)
    st = cfg.strategy
    qw = qs = None
    if isinstance(st, GridTraining):
        ps, bs = generate_training_sets(sys_.domain, st.dx, sys_.eqs, sys_.bcs, np.float64, vi)
        sets = ps + bs
    elif isinstance(st, StochasticTraining):
        pb, bb = get_bounds(sys_.domain, sys_.eqs, sys_.bcs, np.float64, vi, st)
        rng = np.random.default_rng(seed)
        sets = [generate_random_points(st.points, b, np.float64, rng) for b in pb] + \
               [generate_random_points(st.bcs_points, b, np.float64, rng) for b in bb]
    elif isinstance(st, QuasiRandomTraining):
        pb, bb = get_bounds(sys_.domain, sys_.eqs, sys_.bcs, np.float64, vi, st)
        sets = [generate_quasi_random_points(st.points, b, np.float64, seed + i) for i, b in enumerate(pb)] + \
               [generate_quasi_random_points(st.bcs_points, b, np.float64, seed + 100 + i) for i, b in enumerate(bb)]
    else:
        pb, bb = get_bounds(sys_.domain, sys_.eqs, sys_.bcs, np.float64, vi, st)
        sets, qw, qs = [], [], []
        for i, b in enumerate(pb + bb):
            n = st.nodes_per_dim if i < len(pb) else st.bc_nodes_per_dim
            p, w, area = gauss_legendre_box(b, n, np.float64)
            sets.append(p); qw.append(w); qs.append(1.0 / area)
    return sets, qw, qs
