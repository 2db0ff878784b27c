"""Host-side mirror of the reference interface: variable bookkeeping, lowering, training sets, error
behaviour.  CPU only (no engine calls)."""
import numpy as np
import pytest
import sympy as sp

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from neuralpde_jl_b200.strategies import (GridTraining, QuadratureTraining, StochasticTraining, gauss_legendre_box,
                                          generate_training_sets, get_bounds, shard_range)
from oracle import reference as R
from cases import CASES


def _poisson():
    cfg = configs.config2(n=8, width=16, hidden=2)
    return cfg, npde.get_vars(cfg.pde_system.ivs, cfg.pde_system.dvs)


def test_get_vars_and_arguments():
    cfg, vi = _poisson()
    assert vi.depvars == ["u"] and vi.indvars == ["x", "y"] and vi.dict_depvar_input == {"u": ["x", "y"]}
    assert npde.get_argument(cfg.pde_system.eqs, vi) == [["x", "y"]]
    assert npde.get_argument(cfg.pde_system.bcs, vi) == [[0.0, "y"], [1.0, "y"], ["x", 0.0], ["x", 1.0]]
    assert npde.get_variables(cfg.pde_system.bcs, vi) == [["y"], ["y"], ["x"], ["x"]]
    # same answers from the independent restatement in the oracle
    s = cfg.pde_system
    assert R.get_argument(s.bcs, s.ivs, s.dvs) == npde.get_argument(s.bcs, vi)


@pytest.mark.parametrize("name", sorted(CASES))
def test_training_sets_match_oracle(name):
    cfg = CASES[name]()
    if not isinstance(cfg.strategy, GridTraining):
        pytest.skip("grid cases only")
    s = cfg.pde_system
    vi = npde.get_vars(s.ivs, s.dvs)
    ps, bs = generate_training_sets(s.domain, cfg.strategy.dx, s.eqs, s.bcs, np.float64, vi)
    po, bo = R.generate_training_sets(s.domain, cfg.strategy.dx, s.eqs, s.bcs, s.ivs, s.dvs)
    for a, b in zip(ps + bs, po + bo):
        np.testing.assert_array_equal(a, b)


def test_grid_set_layout():
    """(d, N) matrices, first variable fastest, bc constants substituted (reference src/discretize.jl:226-238);
    the PDE set is the full grid because `dif` is never filled (:214-222)."""
    cfg, vi = _poisson()
    s = cfg.pde_system
    ps, bs = generate_training_sets(s.domain, cfg.strategy.dx, s.eqs, s.bcs, np.float32, vi)
    assert ps[0].shape == (2, 64) and ps[0].dtype == np.float32
    assert ps[0][0, 1] > ps[0][0, 0] and ps[0][1, 1] == ps[0][1, 0]          # x fastest
    assert ps[0][0].min() == 0.0 and ps[0][0].max() == 1.0                     # boundary points included
    assert bs[0].shape == (2, 8) and np.all(bs[0][0] == 0.0) and np.all(bs[1][0] == 1.0)
    assert np.all(bs[2][1] == 0.0) and np.all(bs[3][1] == 1.0)


def test_bounds_stochastic_and_quadrature():
    cfg = configs.config3(points=100, bcs_points=10)
    s = cfg.pde_system
    vi = npde.get_vars(s.ivs, s.dvs)
    pb, bb = get_bounds(s.domain, s.eqs, s.bcs, np.float64, vi, cfg.strategy)
    np.testing.assert_allclose(pb[0][0], [0.01, -0.99])
    np.testing.assert_allclose(pb[0][1], [0.99, 0.99])
    np.testing.assert_allclose(bb[0][0], [0.0, -0.99]); np.testing.assert_allclose(bb[0][1], [0.0, 0.99])   # u(0, x)
    np.testing.assert_allclose(bb[1][0], [0.01, -1.0]); np.testing.assert_allclose(bb[1][1], [0.99, -1.0])  # u(t, -1)
    po, bo = R.get_bounds(s.domain, s.eqs, s.bcs, s.ivs, s.dvs, 100)
    for (a0, a1), (b0, b1) in zip(pb + bb, po + bo):
        np.testing.assert_allclose(a0, b0); np.testing.assert_allclose(a1, b1)
    pts, w, area = gauss_legendre_box((np.array([0.0, 2.0, 1.0]), np.array([1.0, 2.0, 3.0])), 5, np.float64)
    assert pts.shape == (3, 25) and np.all(pts[1] == 2.0) and abs(w.sum() - 2.0) < 1e-13 and area == 2.0
    # exact for polynomials of degree <= 9
    assert abs(np.sum(w * pts[0] ** 8 * pts[2] ** 3) - (1 / 9) * (81 - 1) / 4) < 1e-12


def test_lowering_poisson_program():
    cfg, vi = _poisson()
    lt = npde.lower_equation(cfg.pde_system.eqs[0], vi)
    assert [(t.net, t.order, tuple(t.dirs)) for t in lt.taps] == [(0, 2, (0, 0)), (0, 2, (1, 1))]
    assert lt.prog[-1][0] == "sub" and lt.indvars == ["x", "y"] and lt.net_rows == [[0, 1]]
    lb = npde.lower_equation(cfg.pde_system.bcs[0], vi)
    assert [(t.order,) for t in lb.taps] == [(0,)] and lb.indvars == ["x", "y"]


def test_lowering_matches_oracle_residual_numerically():
    """The IR program evaluated on the host equals the oracle's tree walk (taps supplied by the oracle)."""
    import torch
    cfg = CASES["neumann_sin"]()
    s = cfg.pde_system
    vi = npde.get_vars(s.ivs, s.dvs)
    prob = R.Problem(s, cfg.chain_specs(), derivative="exact")
    theta = torch.tensor(cfg.init_params(np.float64))
    X = torch.linspace(0.05, 0.95, 7).reshape(1, -1)
    lt = npde.lower_equation(s.eqs[0], vi)
    dims, acts = cfg.chain_specs()[0]
    taps = [R.exact_tap(X, theta, dims, acts, 0, tuple(t.dirs)).numpy()[0] for t in lt.taps]
    val = []
    for op, a, b, imm in lt.prog:
        v = {"const": lambda: np.full(7, imm), "coord": lambda: X.numpy()[a], "tap": lambda: taps[a],
             "add": lambda: val[a] + val[b], "sub": lambda: val[a] - val[b], "mul": lambda: val[a] * val[b],
             "div": lambda: val[a] / val[b], "neg": lambda: -val[a], "powi": lambda: val[a] ** int(imm),
             "sin": lambda: np.sin(val[a]), "cos": lambda: np.cos(val[a]), "exp": lambda: np.exp(val[a])}[op]()
        val.append(v)
    np.testing.assert_allclose(val[-1], prob.residual(s.eqs[0], X, theta).numpy()[0], rtol=1e-13)


def test_error_behaviour_matches_reference():
    x, y = npde.parameters("x y")
    u = npde.variables("u")
    D3 = npde.Differential(x) ** 3
    vi = npde.get_vars([x, y], [u(x, y)])
    # pure third derivatives lower to one tap (reference stencil src/pinn_types.jl:469-474) ...
    lt = npde.lower_equation(npde.Eq(D3(u(x, y)), 0), vi)
    assert [(t.order, tuple(t.dirs)) for t in lt.taps] == [(3, (0, 0, 0))]
    # ... mixed third derivatives and order 4 (:461-468, and the recursive form :454-460) are refused loudly
    with pytest.raises(npde.LoweringError, match="order 3"):
        npde.lower_equation(npde.Eq(npde.Differential(y)(npde.Differential(x)(npde.Differential(x)(u(x, y)))), 0), vi)
    with pytest.raises(npde.LoweringError, match="order 4"):
        npde.lower_equation(npde.Eq((npde.Differential(x) ** 4)(u(x, y)), 0), vi)
    # trivial bc 0 ~ 0 (reference: ArgumentError for sampling strategies,
    # test/direct_function__trivial_bc_0_0_fails_for_some_training_strategies.jl:41-45)
    with pytest.raises(npde.LoweringError, match="no dependent variable"):
        npde.lower_equation(npde.Eq(0, 0), vi)
    with pytest.raises(npde.LoweringError, match="unknown symbol"):
        npde.lower_equation(npde.Eq(u(x, y), sp.Symbol("q")), vi)
    # empty bcs (reference: solve throws, test/direct_function__empty_boundary_condition_fails_in_solve_phase.jl:15-25)
    sys_ = npde.PDESystem(npde.Eq(u(x, y), 0), [], [npde.In(x, 0, 1), npde.In(y, 0, 1)], [x, y], [u(x, y)])
    disc = npde.PhysicsInformedNN(npde.Chain(npde.Dense(2, 8, "tanh"), npde.Dense(8, 1)), GridTraining(0.5))
    with pytest.raises(ValueError, match="no boundary conditions"):
        npde.symbolic_discretize(sys_, disc)
    with pytest.raises(ValueError, match="custom `derivative`"):
        npde.PhysicsInformedNN(npde.Chain(npde.Dense(2, 8, "tanh"), npde.Dense(8, 1)), GridTraining(0.5), derivative=lambda *a: 0)


def test_theta_layout_is_componentarray_order():
    chain = npde.Chain(npde.Dense(2, 3, "tanh"), npde.Dense(3, 1))
    th = npde.initialparameters(np.random.default_rng(0), chain)
    assert th.size == 2 * 3 + 3 + 3 + 1 and np.all(th[6:9] == 0) and th[12] == 0
    # weight is out x in column-major: the oracle's unpack reads it back
    import torch
    Ws, bs = R.unpack(torch.tensor(th), [2, 3, 1])
    assert Ws[0].shape == (3, 2) and float(Ws[0][1, 0]) == th[1] and float(Ws[0][0, 1]) == th[3]


def test_shard_range_partitions():
    for n, w in ((10, 3), (16384, 8), (5, 8), (0, 2)):
        spans = [shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(hi - lo for lo, hi in spans) - min(hi - lo for lo, hi in spans) <= 1


def test_logging_hooks_noop_and_custom():
    """reference test/qa/qa.jl:32-62: no-op fallback, custom logger methods."""
    npde.logscalar(None, 1.0, "a", 1)
    npde.logvector(None, [1.0], "a", 1)

    class L:
        def __init__(self): self.rows = []
        def log_value(self, name, v, step): self.rows.append((name, v, step))
    lg = L()
    npde.logscalar(lg, 2.0, "loss", 3)
    npde.logvector(lg, [1.0, 2.0], "w", 3)
    assert lg.rows == [("loss", 2.0, 3), ("w/1", 1.0, 3), ("w/2", 2.0, 3)]


def test_gradient_scale_adaptive_loss_rule():
    """bc weights <- a * w + (1 - a) * max|grad pde| / (mean|grad bc_j| + 1e-7), every `reweight_every` iterations
    (reference src/adaptive_losses.jl:100-126)."""
    ada = npde.GradientScaleAdaptiveLoss(3, weight_change_inertia=0.9)
    w = {"pde": np.ones(2), "bc": np.array([1.0, 2.0, 4.0]), "add": np.ones(1)}
    stats = {0: (5.0, 0.1), 1: (7.0, 0.2), 2: (9.0, 0.5), 3: (1.0, 0.25), 4: (3.0, 2.0)}
    ada.update(2, [0.0, 0.0], [0.0, 0.0, 0.0], w, term_grad_stats=lambda i: stats[i])
    np.testing.assert_allclose(w["bc"], [1.0, 2.0, 4.0])                       # not a reweighting iteration
    ada.update(3, [0.0, 0.0], [0.0, 0.0, 0.0], w, term_grad_stats=lambda i: stats[i])
    np.testing.assert_allclose(w["bc"], 0.9 * np.array([1.0, 2.0, 4.0]) + 0.1 * 7.0 / (np.array([0.5, 0.25, 2.0]) + 1e-7))
    np.testing.assert_allclose(w["pde"], [1.0, 1.0])


def test_minimax_adaptive_loss_uses_the_optimiser_rule():
    """reference src/adaptive_losses.jl:192-195, :223-237: Optimisers.update!(setup(Adam(lr)), weights, -losses).  The first
    Adam step moves every weight by ~lr whatever the loss magnitude (m / sqrt(v) = sign), the second follows the moment
    recursion; Descent(lr) is plain ascent."""
    ada = npde.MiniMaxAdaptiveLoss(2)
    assert ada.pde_max_optimiser.lr == 1e-4 and ada.bc_max_optimiser.lr == 0.5
    w = {"pde": np.ones(1), "bc": np.array([1.0, 3.0]), "add": np.ones(1)}
    ada.update(1, [10.0], [1e-3, 7.0], w)
    np.testing.assert_allclose(w["bc"], [1.0, 3.0]); np.testing.assert_allclose(w["pde"], [1.0])
    ada.update(2, [10.0], [1e-3, 7.0], w)
    np.testing.assert_allclose(w["pde"], [1.0 + 1e-4], rtol=1e-9)
    np.testing.assert_allclose(w["bc"], [1.5, 3.5], rtol=1e-4)          # eps = 1e-8 against |g| = 1e-3 shows at 1e-5
    g1, g2 = np.array([-1e-3, -7.0]), np.array([-2e-3, -1.0])
    ada.update(4, [10.0], -g2, w)
    m = 0.9 * (0.1 * g1) + 0.1 * g2
    v = 0.999 * (0.001 * g1 ** 2) + 0.001 * g2 ** 2
    step = 0.5 * (m / (1 - 0.9 ** 2)) / (np.sqrt(v / (1 - 0.999 ** 2)) + 1e-8)
    np.testing.assert_allclose(w["bc"], np.array([1.0, 3.0]) + 0.5 * (0.1 * -g1 / 0.1) / (np.sqrt(0.001 * g1 ** 2 / 0.001) + 1e-8) - step,
                               rtol=1e-12)
    asc = npde.MiniMaxAdaptiveLoss(1, pde_max_optimiser=npde.Descent(0.1), bc_max_optimiser=npde.Descent(0.5))
    w2 = {"pde": np.ones(1), "bc": np.ones(2), "add": np.ones(1)}
    asc.update(1, [2.0], [4.0, 6.0], w2)
    np.testing.assert_allclose(w2["pde"], [1.2]); np.testing.assert_allclose(w2["bc"], [3.0, 4.0])


def test_softadapt_and_relobralo_rules():
    """reference src/adaptive_losses.jl:330-361 (SoftAdapt: softmax(alpha * relative change since the last reweighting) * N)
    and :461-488 (ReLoBRaLo: softmax(alpha * L / L_ref) * N with L_ref = previous reweighting w.p. beta, else the first call)."""
    def softmax(x):
        e = np.exp(x - x.max()); return e / e.sum()
    sa = npde.SoftAdaptAdaptiveLoss(2, alpha=0.1)
    w = {"pde": np.ones(1), "bc": np.ones(2), "add": np.ones(1)}
    L1, L2, L3 = np.array([4.0, 1.0, 0.5]), np.array([2.0, 1.5, 0.5]), np.array([1.0, 3.0, 0.25])
    sa.update(1, L1[:1], L1[1:], w)                       # seeds prev, no reweighting
    np.testing.assert_allclose(np.r_[w["pde"], w["bc"]], 1.0)
    sa.update(2, L2[:1], L2[1:], w)
    np.testing.assert_allclose(np.r_[w["pde"], w["bc"]], softmax(0.1 * (L2 - L1) / (L1 + 1e-8)) * 3, rtol=1e-13)
    sa.update(3, L3[:1], L3[1:], w)                       # not a reweighting iteration: prev stays L2
    sa.update(4, L3[:1], L3[1:], w)
    np.testing.assert_allclose(np.r_[w["pde"], w["bc"]], softmax(0.1 * (L3 - L2) / (L2 + 1e-8)) * 3, rtol=1e-13)
    assert abs(w["pde"].sum() + w["bc"].sum() - 3.0) < 1e-12
    for beta, ref in ((0.0, L1), (1.0, L2)):              # beta = 0: always the initial losses; beta = 1: always the previous
        rl = npde.ReLoBRaLoAdaptiveLoss(1, alpha=1.0, beta=beta, seed=0)
        w = {"pde": np.ones(1), "bc": np.ones(2), "add": np.ones(1)}
        rl.update(1, L1[:1], L1[1:], w)
        np.testing.assert_allclose(np.r_[w["pde"], w["bc"]], softmax(L1 / (L1 + 1e-8)) * 3, rtol=1e-13)
        rl.update(2, L2[:1], L2[1:], w)
        rl.update(3, L3[:1], L3[1:], w)
        np.testing.assert_allclose(np.r_[w["pde"], w["bc"]], softmax(L3 / (ref + 1e-8)) * 3, rtol=1e-13)
    rl = npde.ReLoBRaLoAdaptiveLoss(1, beta=0.5, seed=123)
    w = {"pde": np.ones(1), "bc": np.ones(2), "add": np.ones(1)}
    draws = []
    for it, L in enumerate([L1, L2, L3, L1, L2, L3, L1, L2], start=1):
        rl.update(it, L[:1], L[1:], w); draws.append(rl.last_use_prev)
    assert any(draws) and not all(draws)                  # both references occur at beta = 0.5


def _run_ir(prog, rows, taps):
    """Evaluate a lowered residual program on the host (rows: point-matrix rows incl. hoisted ones)."""
    val = []
    for op, a, b, imm in prog:
        f = {"const": lambda: np.full(rows.shape[1], imm), "coord": lambda: rows[a], "tap": lambda: taps[a],
             "add": lambda: val[a] + val[b], "sub": lambda: val[a] - val[b], "mul": lambda: val[a] * val[b],
             "div": lambda: val[a] / val[b], "neg": lambda: -val[a], "powi": lambda: val[a] ** int(imm),
             "pow": lambda: val[a] ** val[b], "sin": lambda: np.sin(val[a]), "cos": lambda: np.cos(val[a]),
             "exp": lambda: np.exp(val[a]), "log": lambda: np.log(val[a]), "tanh": lambda: np.tanh(val[a]),
             "sqrt": lambda: np.sqrt(val[a]), "abs": lambda: np.abs(val[a])}[op]
        val.append(f())
    return val[-1]


@pytest.mark.parametrize("name", ["cfg2_small", "neumann_sin", "mixed", "cfg3_small"])
def test_hoisted_and_plain_lowering_agree(name):
    """Hoisting coordinate-only subexpressions into extra point rows (host-evaluated once per point set) must not change
    the residual: the device-side sampler path lowers without hoisting, every other path with it."""
    cfg = CASES[name]()
    s = cfg.pde_system
    vi = npde.get_vars(s.ivs, s.dvs)
    rng = np.random.default_rng(3)
    for eq in list(s.eqs) + list(s.bcs):
        plain = npde.lower_equation(eq, vi, hoist=False)
        hoisted = npde.lower_equation(eq, vi, hoist=True)
        assert len(plain.taps) == len(hoisted.taps)
        d = len(vi.indvars) if hasattr(vi, "indvars") else 2
        X = rng.uniform(0.1, 0.9, size=(plain.augment(np.zeros((d, 1))).shape[0], 9))
        taps = [rng.standard_normal(9) for _ in plain.taps]
        r_plain = _run_ir(plain.prog, plain.augment(X), taps)
        r_hoist = _run_ir(hoisted.prog, hoisted.augment(X), taps)
        np.testing.assert_allclose(r_hoist, r_plain, rtol=1e-13, atol=1e-14)
        assert plain.augment(X).shape[0] == X.shape[0]                  # no extra rows without hoisting


def test_bayesian_pinn_wraps_a_physics_informed_nn_and_validates():
    """BayesianPINN(args...; dataset, kwargs...) forwards to PhysicsInformedNN (reference src/pinn_types.jl:231-245); the
    log-likelihood form exists for GridTraining only (src/training_strategies.jl:50-113)."""
    from neuralpde_jl_b200 import configs
    cfg = configs.config2(n=8, width=8, hidden=2)
    b = npde.BayesianPINN(cfg.chains[0], cfg.strategy, param_estim=False)
    assert isinstance(b.pinn, npde.PhysicsInformedNN) and tuple(b.dataset) == (None, None) and b.strategy is cfg.strategy
    with pytest.raises(ValueError, match="GridTraining only"):
        npde.symbolic_discretize(cfg.pde_system, npde.BayesianPINN(cfg.chains[0], npde.StochasticTraining(16)))
    with pytest.raises(ValueError, match="dataset points"):
        npde.symbolic_discretize(cfg.pde_system, npde.BayesianPINN(cfg.chains[0], cfg.strategy, dataset=(np.zeros((3, 3)), None)))


def test_device_loop_breaks_at_reweighting_iterations():
    """solve(device_loop=True) with an adaptive loss: the device-resident loop runs up to the iteration before each
    reweighting, the reweighting iteration evaluates the term losses at the current theta, updates the weights and takes
    its step with the NEW weights (order of src/discretize.jl:574-588).  Driven with a stub engine (no GPU)."""
    from neuralpde_jl_b200.pinn import OptimizationFunction, OptimizationProblem, solve

    class StubEngine:
        def __init__(self):
            self.log, self.steps = [], 0
        def adam_begin(self, th, *a): self.log.append(("begin",))
        def adam_theta(self): return np.full(3, float(self.steps))
        def loss_grad_host(self, th, w, want_grad):
            self.log.append(("loss", self.steps, tuple(np.round(w, 6))))
            return 1.0, np.array([2.0 + self.steps, 1.0, 3.0]), None
        def term_grad_stats_host(self, i, th): return (1.0, 1.0)
        def adam_iterate(self, n, w):
            self.log.append(("iterate", n, tuple(np.round(w, 6))))
            self.steps += n
            return 0.5, np.zeros(3)

    class Rep:
        pass
    rep = Rep()
    rep.strategy = GridTraining(0.1); rep.engine = StubEngine()
    rep.adaloss = npde.MiniMaxAdaptiveLoss(4, pde_max_optimiser=npde.Descent(0.1), bc_max_optimiser=npde.Descent(0.5))
    rep.weights = {"pde": np.ones(1), "bc": np.ones(2), "add": np.ones(1)}
    rep.iteration = [0]; rep.eqs = [0]; rep.bcs = [0, 1]; rep.additional_loss = None
    prob = OptimizationProblem(OptimizationFunction(None, None), np.zeros(3), None, rep)
    sol = solve(prob, npde.Adam(0.01), maxiters=10, device_loop=True, chunk=50)
    log = rep.engine.log
    assert sol.iterations == 10 and rep.iteration[0] == 10
    its = [e for e in log if e[0] == "iterate"]
    # 3 iterations with unit weights, reweight at iteration 4 (theta after 3 steps), 1 step with the new weights, 3 more,
    # reweight at 8 (theta after 7 steps), 1 step, then the remaining 2
    assert [e[1] for e in its] == [3, 1, 3, 1, 2]
    assert its[0][2] == (1.0, 1.0, 1.0)
    w4 = (1.0 + 0.1 * 5.0, 1.0 + 0.5 * 1.0, 1.0 + 0.5 * 3.0)            # losses at steps = 3: [5, 1, 3]
    assert its[1][2] == w4 and its[2][2] == w4
    w8 = (w4[0] + 0.1 * 9.0, w4[1] + 0.5, w4[2] + 1.5)                  # losses at steps = 7: [9, 1, 3]
    assert its[3][2] == w8 and its[4][2] == w8
    losses = [e for e in log if e[0] == "loss"]
    assert [e[1] for e in losses] == [0, 3, 7]                          # the seeding call + the two reweightings
