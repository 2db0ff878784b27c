"""Size-independent properties at BASELINE.json's full config-2 size (128^2 points, 4x64 tanh MLP)."""
import numpy as np
import pytest

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from neuralpde_jl_b200.strategies import shard_range
from helpers import rel

pytestmark = pytest.mark.gpu


def _rep(mode="ffma", dtype=np.float32):
    cfg = configs.config2()
    return cfg, npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode))


def test_gradient_is_deterministic_and_linear_in_weights():
    cfg, rep = _rep()
    th = rep.flat_init_params
    t1, _, g1 = rep.engine.loss_grad_host(th, None, True)
    t2, _, g2 = rep.engine.loss_grad_host(th, None, True)
    assert t1 == t2 and np.array_equal(g1, g2)              # fixed-order reduction: bitwise reproducible
    _, _, ga = rep.engine.loss_grad_host(th, np.array([1.0, 0, 0, 0, 0]), True)
    _, _, gb = rep.engine.loss_grad_host(th, np.array([0, 1.0, 1.0, 1.0, 1.0]), True)
    assert rel(ga + gb, g1) < 1e-6


def test_sharded_sums_equal_full():
    """Two half-size shards with n_global set reproduce the full loss and gradient (what the allreduce adds)."""
    cfg, rep = _rep(dtype=np.float64)
    th = rep.flat_init_params
    total, terms, grad = rep.engine.loss_grad_host(th, None, True)
    acc_t, acc_g = np.zeros_like(terms), np.zeros_like(grad)
    for r in range(2):
        cfg2, rr = _rep(dtype=np.float64)
        for i, s in enumerate(rep.point_sets[:5]):
            lo, hi = shard_range(s.shape[1], r, 2)
            rr.set_points(i, s[:, lo:hi], n_global=s.shape[1])
        _, t_r, g_r = rr.engine.loss_grad_host(th, None, True)
        acc_t += t_r; acc_g += g_r
    np.testing.assert_allclose(acc_t, terms, rtol=1e-12)
    assert rel(acc_g, grad) < 1e-12


def test_tc_split_agrees_with_ffma_at_full_size():
    cfg, rf = _rep("ffma")
    _, rs = _rep("tc_split")
    th = rf.flat_init_params
    tf, termsf, gf = rf.engine.loss_grad_host(th, None, True)
    ts, termss, gs = rs.engine.loss_grad_host(th, None, True)
    assert abs(ts - tf) <= 1e-5 * abs(tf)
    assert rel(gs, gf) < 1e-2
    # loss-only call returns the same total and no gradient
    ts2, _, g2 = rs.engine.loss_grad_host(th, None, False)
    assert g2 is None and abs(ts2 - ts) <= 1e-6 * abs(ts)


def test_phi_prediction_and_adam_step():
    """phi(x, θ) keeps working on the engine; one Adam step through discretize/solve is finite
    (reference src/precompilation.jl:10-24, test/Interface/precompile_workload.jl:17-28)."""
    cfg = configs.config2(n=16, width=16, hidden=2)
    prob = npde.discretize(cfg.pde_system, cfg.discretization(dtype=np.float32))
    res = npde.solve(prob, npde.Adam(0.01), maxiters=3)
    assert np.isfinite(res.objective) and res.u.shape == prob.u0.shape
    phi = prob.representation.phi
    out = phi(np.array([[0.25, 0.5], [0.75, 0.5]]), res.u)
    assert out.shape == (1, 2) and np.all(np.isfinite(out))
    from oracle import reference as R
    import torch
    ref = R.phi(torch.tensor([[0.25, 0.5], [0.75, 0.5]]), torch.tensor(res.u.astype(np.float64)), *cfg.chain_specs()[0]).numpy()
    np.testing.assert_allclose(out, ref, rtol=1e-5, atol=1e-6)


def test_device_adam_loop_matches_host_adam():
    """pinn_adam_iterate (Adam fused into the gradient reduction, theta resident on the device) follows the same
    trajectory as the host Adam loop that calls pinn_loss_grad every iteration."""
    cfg = configs.config2(n=16, width=16, hidden=2)
    prob_h = npde.discretize(cfg.pde_system, cfg.discretization(dtype=np.float64))
    prob_d = npde.discretize(cfg.pde_system, cfg.discretization(dtype=np.float64))
    res_h = npde.solve(prob_h, npde.Adam(0.01), maxiters=7)
    res_d = npde.solve(prob_d, npde.Adam(0.01), maxiters=7, device_loop=True, chunk=3)
    assert res_d.iterations == 7
    np.testing.assert_allclose(res_d.u, res_h.u, rtol=1e-9, atol=1e-12)
    assert abs(res_d.objective - res_h.objective) <= 1e-9 * abs(res_h.objective)
    # fp32 + tensor-core mode: a few steps reduce the loss
    cfg2 = configs.config2(n=32, width=32, hidden=3)
    prob_t = npde.discretize(cfg2.pde_system, cfg2.discretization(dtype=np.float32, mode="tc_split"))
    l0 = prob_t.f.f(prob_t.u0, None)
    res_t = npde.solve(prob_t, npde.Adam(0.003), maxiters=40, device_loop=True, chunk=20)
    assert np.isfinite(res_t.objective) and res_t.objective < l0


def test_device_loop_rejects_resampled_sets():
    cfg = configs.config3(points=256, bcs_points=32, width=16, hidden=2)
    prob = npde.discretize(cfg.pde_system, cfg.discretization(dtype=np.float32))
    with pytest.raises(ValueError, match="point sets that live on the device"):
        npde.solve(prob, npde.Adam(0.01), maxiters=2, device_loop=True)


def test_device_sampler_bounds_determinism_and_oracle_parity():
    """StochasticTraining with the device-side sampler (pinn_set_sampler / pinn_resample; the reference draws on the host and
    uploads every call, src/training_strategies.jl:271-282): points stay inside the reference's bounds (get_bounds,
    src/discretize.jl:299-324), boundary sets keep their constant coordinate, a second engine with the same seed draws the
    same sequence, each call draws fresh points, and the loss at the drawn points equals the float64 oracle's."""
    from helpers import oracle_eval
    from neuralpde_jl_b200.strategies import get_bounds
    from neuralpde_jl_b200.symbolic import get_vars

    def make():
        cfg = configs.config3(points=1500, bcs_points=200, width=16, hidden=2)
        cfg.strategy.device_sampler = True
        return cfg, npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float64))

    cfg, rep = make()
    sys_ = cfg.pde_system
    vi = get_vars(sys_.ivs, sys_.dvs)
    pb, bb = get_bounds(sys_.domain, sys_.eqs, sys_.bcs, np.float64, vi, cfg.strategy)
    counts = [1500, 200, 200, 200]
    th = rep.flat_init_params
    total = rep.loss_functions.full_loss_function(th)
    pts = [rep.engine.get_points_host(i, n) for i, n in enumerate(counts)]
    for p_, b in zip(pts, pb + bb):
        for r, (lo, hi) in enumerate(zip(*b)):
            assert p_[r].min() >= lo and p_[r].max() <= hi
            if lo == hi:
                assert np.all(p_[r] == lo)
            else:       # uniform: mean within 5 sigma, both halves populated
                assert abs(p_[r].mean() - 0.5 * (lo + hi)) < 5 * (hi - lo) / np.sqrt(12 * p_.shape[1])
    L, T, G = oracle_eval(cfg, th.astype(np.float64), "exact", pts)
    assert abs(total - L) <= 1e-10 * abs(L)
    total2 = rep.loss_functions.full_loss_function(th)          # second call: fresh draw
    pts2 = rep.engine.get_points_host(0, 1500)
    assert not np.array_equal(pts2, pts[0]) and total2 != total
    _, rep_b = make()
    t_b = rep_b.loss_functions.full_loss_function(th)
    assert t_b == total and np.array_equal(rep_b.engine.get_points_host(0, 1500), pts[0])
    # device-resident Adam loop: a fresh sample every step, no host round trip
    res = npde.solve(npde.discretize(cfg.pde_system, make()[0].discretization(dtype=np.float64)), npde.Adam(1e-3), maxiters=25,
                     device_loop=True)
    assert np.isfinite(res.objective) and np.all(np.isfinite(res.u))


@pytest.mark.parametrize("mode,dtype,tol", [("ffma", np.float64, 1e-13), ("tc_split", np.float32, 2e-6), ("tc_bf16", np.float32, 2e-6)])
def test_in_kernel_tail_matches_separate_reduce_kernel(monkeypatch, mode, dtype, tol):
    """The gradient reduction in the fused kernel's tail (one launch per step) gives the round-1 sequence's result
    (fused kernel -> reduce_kernel, selected with PINN_B200_TAIL=0) up to summation order, and launches once."""
    cfg = configs.config2(n=48, width=32, hidden=3)
    rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode))
    th = rep.flat_init_params
    n0 = rep.engine.launch_count()
    t1, terms1, g1 = rep.engine.loss_grad_host(th, None, True)
    assert rep.engine.launch_count() - n0 == 1
    monkeypatch.setenv("PINN_B200_TAIL", "0")
    rep0 = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode))
    monkeypatch.delenv("PINN_B200_TAIL")
    n0 = rep0.engine.launch_count()
    t0, terms0, g0 = rep0.engine.loss_grad_host(th, None, True)
    assert rep0.engine.launch_count() - n0 == 2
    assert abs(t1 - t0) <= tol * abs(t0) and rel(g1, g0) < 10 * tol
    np.testing.assert_allclose(terms1, terms0, rtol=10 * tol)
    # many steps through one handle: the self-resetting grid barrier and the step counter stay consistent
    for _ in range(50):
        t2, _, g2 = rep.engine.loss_grad_host(th, None, True)
    if mode == "ffma":        # the FFMA kernel's partials have one writer per entry: bitwise reproducible
        assert t2 == t1 and np.array_equal(g2, g1)
    else:                     # the tcgen05 kernels combine a few warps' sums per entry with atomics inside a CTA
        assert abs(t2 - t1) <= 1e-6 * abs(t1) and rel(g2, g1) < 1e-6


def test_adam_graph_replay_matches_uncaptured_loop(monkeypatch):
    """pinn_adam_iterate captures its iterations into a CUDA graph (step counter / bias correction on the device);
    replaying it gives the same trajectory as the uncaptured launch loop (PINN_B200_NO_GRAPH=1)."""
    cfg = configs.config2(n=24, width=16, hidden=2)
    def run():
        rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float64))
        rep.engine.adam_begin(rep.flat_init_params, 0.01)
        losses = [rep.engine.adam_iterate(4)[0] for _ in range(3)]      # 2nd and 3rd call replay the graph
        return rep.engine.adam_theta(), losses
    th_g, l_g = run()
    monkeypatch.setenv("PINN_B200_NO_GRAPH", "1")
    th_n, l_n = run()
    assert np.array_equal(th_g, th_n) and l_g == l_n
    assert l_g[2] < l_g[0]
