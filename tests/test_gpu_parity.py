"""GPU parity: the CUDA path (through the C ABI) against the float64 oracle on the same
seeded inputs.  Tolerances: fp64 engine 1e-9 (same algorithm, different summation order);
fp32 engine loss rtol 1e-5 (BASELINE.json north_star), gradient relative L2 error 2e-4."""
import numpy as np
import pytest

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from helpers import engine_eval, oracle_eval, rel

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,kw", [("cfg1", {}), ("cfg2", dict(n=24, width=16, hidden=2)), ("cfg2", dict(n=40))])
def test_fp64_engine_matches_oracle(name, kw):
    cfg = configs.ALL[name](**kw)
    rep, total, terms, grad = engine_eval(cfg, np.float64)
    L, T, G = oracle_eval(cfg, rep.flat_init_params.astype(np.float64))
    assert abs(total - L) <= 1e-10 * abs(L)
    np.testing.assert_allclose(terms, T, rtol=1e-10)
    assert rel(grad, G) < 1e-9


@pytest.mark.parametrize("name,kw", [("cfg1", {}), ("cfg2", dict(n=24, width=16, hidden=2)), ("cfg2", {})])
def test_fp32_engine_loss_rtol_1e5(name, kw):
    cfg = configs.ALL[name](**kw)
    rep, total, terms, grad = engine_eval(cfg, np.float32)
    L, T, G = oracle_eval(cfg, rep.flat_init_params.astype(np.float64))
    assert abs(total - L) <= 1e-5 * abs(L), (total, L)
    np.testing.assert_allclose(terms, T, rtol=1e-5)
    assert rel(grad, G) < 2e-4


def test_fp32_engine_vs_reference_fd_semantics():
    """Against the reference-faithful finite-difference path in float64 (what NeuralPDE.jl
    computes with its default eltype)."""
    cfg = configs.config2(n=32)
    rep, total, terms, grad = engine_eval(cfg, np.float32)
    L, T, G = oracle_eval(cfg, rep.flat_init_params.astype(np.float64), derivative="fd")
    assert abs(total - L) <= 1e-5 * abs(L)
    assert rel(grad, G) < 2e-4


def test_residual_probe_matches_oracle():
    cfg = configs.config2(n=20, width=16, hidden=2)
    rep, *_ = engine_eval(cfg, np.float64, want_grad=False)
    from oracle import reference as R
    import torch
    prob = R.Problem(cfg.pde_system, cfg.chain_specs(), derivative="exact")
    pts = rep.point_sets[0]
    r = rep.loss_functions.datafree_pde_loss_functions[0](pts, rep.flat_init_params)
    ro = prob.residual(cfg.pde_system.eqs[0], torch.as_tensor(pts), torch.as_tensor(rep.flat_init_params)).numpy()
    np.testing.assert_allclose(r, ro, rtol=1e-9, atol=1e-11)


def test_term_weights_and_loss_only():
    cfg = configs.config2(n=16, width=16, hidden=2)
    disc = cfg.discretization(dtype=np.float64)
    rep = npde.symbolic_discretize(cfg.pde_system, disc)
    th = rep.flat_init_params
    w = np.array([2.0, 0.5, 1.0, 3.0, 0.25])
    total, terms, grad = rep.engine.loss_grad_host(th, w, True)
    assert abs(total - float(np.dot(w, terms))) < 1e-12 * abs(total)
    total2, terms2, g2 = rep.engine.loss_grad_host(th, w, False)
    assert g2 is None and abs(total2 - total) < 1e-13 * abs(total)
    # gradient is linear in the weights
    _, _, g_a = rep.engine.loss_grad_host(th, np.array([1.0, 0, 0, 0, 0]), True)
    _, _, g_b = rep.engine.loss_grad_host(th, np.array([0, 1.0, 1.0, 1.0, 1.0]), True)
    _, _, g_ab = rep.engine.loss_grad_host(th, np.ones(5), True)
    assert rel(g_a + g_b, g_ab) < 1e-12


def test_term_grad_stats_and_gradient_scale_adaptive_loss():
    """pinn_term_grad_stats == max / mean of |gradient of one unweighted term| (reference: Zygote.gradient per term,
    src/adaptive_losses.jl:107-116); GradientScaleAdaptiveLoss moves the boundary weights by the stated EMA rule."""
    cfg = configs.config2(n=16, width=16, hidden=2)
    rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float64))
    th = rep.flat_init_params
    n_terms = rep.engine.n_terms
    stats = []
    for i in range(n_terms):
        w = np.zeros(n_terms); w[i] = 1.0
        _, _, g = rep.engine.loss_grad_host(th, w, True)
        mx, mn = rep.engine.term_grad_stats_host(i, th)
        assert abs(mx - np.max(np.abs(g))) <= 1e-13 * mx and abs(mn - np.mean(np.abs(g))) <= 1e-12 * mn
        stats.append((mx, mn))
    ada = npde.GradientScaleAdaptiveLoss(2, weight_change_inertia=0.5)
    disc = cfg.discretization(dtype=np.float64)
    disc.adaptive_loss = ada
    rep2 = npde.symbolic_discretize(cfg.pde_system, disc)
    f = rep2.loss_functions.full_loss_function
    terms = rep.engine.loss_grad_host(th, None, False)[1]
    total1 = f(th)                           # iteration 1: no reweighting
    assert abs(total1 - float(np.sum(terms))) <= 1e-12 * abs(total1)
    # iteration 2 reweights BEFORE forming the weighted sum (src/discretize.jl:574-588): the returned loss already uses
    # the new boundary weights
    total2 = f(th)
    expected = 0.5 * 1.0 + 0.5 * stats[0][0] / (np.array([s_[1] for s_ in stats[1:]]) + 1e-7)
    assert abs(total2 - (terms[0] + float(np.dot(expected, terms[1:])))) <= 1e-10 * abs(total2)
    total3 = f(th)                           # iteration 3: same weights, no reweighting
    assert abs(total3 - total2) <= 1e-12 * abs(total2)
    # the gradient of a reweighting iteration uses the new weights too
    wfull = np.concatenate([[1.0], expected])
    _, _, g_expected = rep.engine.loss_grad_host(th, wfull, True)
    disc_b = cfg.discretization(dtype=np.float64)
    disc_b.adaptive_loss = npde.GradientScaleAdaptiveLoss(1, weight_change_inertia=0.5)
    rep_b = npde.symbolic_discretize(cfg.pde_system, disc_b)
    tot_b, g_b = rep_b.loss_functions.full_loss_gradient(th)          # iteration 1 is already a reweighting iteration
    assert abs(tot_b - total2) <= 1e-10 * abs(total2) and rel(g_b, g_expected) < 1e-12


def test_theta_layout_hand_computed_network_on_the_engine():
    """The engine reads theta in the reference's ComponentArray order (weight out x in column-major, then bias, layer by
    layer): phi of a hand-computed 2 -> 2 -> 1 network, and d(loss)/d(theta) entry by entry against central differences
    of the hand formula (tests/test_oracle_pinning.py holds the same network for the oracle)."""
    from test_oracle_pinning import HAND_THETA, hand_phi
    chain = npde.Chain(npde.Dense(2, 2, "tanh"), npde.Dense(2, 1))
    phi = npde.Phi(chain, 0, HAND_THETA.size, np.float64)
    pts = np.array([[0.2, -1.0, 0.7], [0.9, 0.4, -0.3]])
    want = np.array([hand_phi(x, y) for x, y in pts.T])
    np.testing.assert_allclose(phi(pts, HAND_THETA)[0], want, rtol=1e-13)
    # loss = mean(abs2, u(x, y) - 0.5) over the three points, through the full discretize path
    x, y = npde.parameters("x y")
    u = npde.variables("u")
    sys_ = npde.PDESystem(npde.Eq(u(x, y), 0.5), [npde.Eq(u(0, y), 0.0)], [npde.In(x, 0.0, 1.0), npde.In(y, 0.0, 1.0)], [x, y], [u(x, y)])
    rep = npde.symbolic_discretize(sys_, npde.PhysicsInformedNN(chain, npde.GridTraining(0.5), init_params=HAND_THETA))
    rep.set_points(0, pts)
    _, terms, grad = rep.engine.loss_grad_host(HAND_THETA, np.array([1.0, 0.0]), True)
    loss = lambda th: float(np.mean([(hand_phi(a, b, th) - 0.5) ** 2 for a, b in pts.T]))     # noqa: E731
    assert abs(terms[0] - loss(HAND_THETA)) < 1e-14
    for i in range(HAND_THETA.size):
        e = np.zeros_like(HAND_THETA); e[i] = 1e-6
        fd = (loss(HAND_THETA + e) - loss(HAND_THETA - e)) / 2e-6
        assert abs(grad[i] - fd) < 1e-8, (i, grad[i], fd)


def test_bayesian_pinn_loglikelihood_and_gradient():
    """BayesianPINN's objective (reference src/discretize.jl:653-757, src/training_strategies.jl:115-128): the weighted sum
    of logpdf(MvNormal(residual, sigma^2 I), 0) over the Grid sets, plus logpdf(Normal(0, stdextra), additional_loss).
    Checked against the same formula evaluated with the float64 oracle's residuals (autograd for the gradient)."""
    import torch
    from oracle import reference as R
    cfg = configs.config2(n=12, width=16, hidden=2)
    rng = np.random.default_rng(2)
    X = rng.random((2, 40))
    data = npde.DataLoss("u", X, np.sin(np.pi * X[0]) * np.sin(np.pi * X[1]) / (2 * np.pi ** 2))
    theta = cfg.init_params(np.float64, seed=3)
    disc = npde.BayesianPINN(cfg.chains[0], cfg.strategy, init_params=theta, additional_loss=data)
    rep = npde.symbolic_discretize(cfg.pde_system, disc)
    allstd = [[0.05], [0.1, 0.2, 0.3, 0.4], 0.7]
    ll, g = rep.loss_functions.full_loss_gradient(theta, allstd)
    assert abs(rep.loss_functions.full_loss_function(theta, allstd) - ll) <= 1e-12 * abs(ll)
    # oracle: residual vectors -> the reference's formula
    sys_ = cfg.pde_system
    prob = R.Problem(sys_, cfg.chain_specs(), derivative="exact")
    ps, bs = R.generate_training_sets(sys_.domain, cfg.strategy.dx, sys_.eqs, sys_.bcs, sys_.ivs, sys_.dvs)
    th = torch.tensor(theta, requires_grad=True)
    def logpdf0(r, s):          # logpdf(MvNormal(r, s^2 I), 0)
        n = r.numel()
        return -0.5 * n * np.log(2 * np.pi) - n * np.log(s) - (r ** 2).sum() / (2 * s ** 2)
    ll_pde = sum(logpdf0(prob.residual(eq, torch.as_tensor(c), th), s) for eq, c, s in zip(sys_.eqs, ps, allstd[0]))
    ll_bc = sum(logpdf0(prob.residual(eq, torch.as_tensor(c), th), s) for eq, c, s in zip(sys_.bcs, bs, allstd[1]))
    A = prob.data_loss(data.depvar, data.points, data.values)(th)
    ll_add = -np.log(allstd[2] * np.sqrt(2 * np.pi)) - A ** 2 / (2 * allstd[2] ** 2)
    want = 1.0 * ll_pde + 4.0 * ll_bc + ll_add        # unit weights: every weight of a group scales the group SUM (:732-738)
    want.backward()
    assert abs(ll - float(want)) <= 1e-10 * abs(float(want)), (ll, float(want))
    assert rel(g, th.grad.numpy()) < 1e-9
