"""The `lower` pass of ext/NeuralPDEB200Ext.jl, exercised through its Python twin (tests/julia_expr.py): the generated loss
functions of the BASELINE configurations (tests/golden/genfn/*.jl, the Expr text build_symbolic_loss_function produces,
src/discretize.jl:28-152) lower to the same residual IR `lowering.py` emits from the symbolic equation -- same taps, same
network-input rows, same values and parameter dependence at random points."""
import os

import numpy as np
import pytest

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from julia_expr import Sym, lower_loss_function, parse

HERE = os.path.dirname(os.path.abspath(__file__))


def _text(name):
    with open(os.path.join(HERE, "golden", "genfn", name + ".jl"), encoding="utf-8") as f:
        return f.read()


def _run(prog, rows, taps, params):
    val = []
    n = rows.shape[1]
    for op, a, b, imm in prog:
        f = {"const": lambda: np.full(n, imm), "coord": lambda: rows[a], "tap": lambda: taps[a], "param": lambda: np.full(n, params[a]),
             "add": lambda: val[a] + val[b], "sub": lambda: val[a] - val[b], "mul": lambda: val[a] * val[b],
             "div": lambda: val[a] / val[b], "neg": lambda: -val[a], "powi": lambda: val[a] ** int(imm),
             "pow": lambda: val[a] ** val[b], "sin": lambda: np.sin(val[a]), "cos": lambda: np.cos(val[a]),
             "exp": lambda: np.exp(val[a]), "log": lambda: np.log(val[a]), "tanh": lambda: np.tanh(val[a]),
             "sqrt": lambda: np.sqrt(val[a]), "abs": lambda: np.abs(val[a])}[op]
        val.append(f())
    return val[-1]


def test_parser_reads_the_appendix_a_function():
    fn = parse(_text("cfg2_pde"))
    assert fn[0] == "->" and [a.name for a in fn[1][1]] == ["cord", "θ", "phi", "derivative", "integral", "u", "p"]
    let = fn[2][1][0][1][0]
    assert let[0] == "let" and [s.name for s in let[1][1][1]] == ["x", "y"]
    loss = let[2][1][-1]
    assert loss[0] == "call" and loss[1] == Sym(".-")
    lhs = loss[2][0]
    assert lhs[0] == "." and lhs[1] == Sym("+") and lhs[2][0][1] == Sym("derivative")
    eps = lhs[2][0][2][3]
    assert eps[0] == "vect" and eps[1][0][1] == [0.0001220703125, 0.0]


CASES = [
    # fixture, config, which equation, kwargs of the lowering
    ("cfg2_pde", lambda: configs.config2(), ("eqs", 0), {}),
    ("cfg2_bc", lambda: configs.config2(), ("bcs", 0), {}),
    ("cfg3_pde", lambda: configs.config3(points=64, bcs_points=8), ("eqs", 0), {}),
    ("cfg3_ic", lambda: configs.config3(points=64, bcs_points=8), ("bcs", 0), {}),
    ("cfg4_momentum_x", lambda: configs.config4(nodes=2, bc_nodes=2, width=16, hidden=2), ("eqs", 0), {}),
    ("cfg4_continuity", lambda: configs.config4(nodes=2, bc_nodes=2, width=16, hidden=2), ("eqs", 3), {}),
    ("cfg5_pde", lambda: configs.config5(points=16, bcs_points=8, n_obs=4, width=16, hidden=2), ("eqs", 0), {"param_estim": True}),
    ("cfg5_pde_fixed_p", lambda: configs.config5(points=16, bcs_points=8, n_obs=4, width=16, hidden=2), ("eqs", 0),
     {"param_estim": False, "default_p": [0.5]}),
]


@pytest.mark.parametrize("fixture,make,which,kw", CASES, ids=[c[0] for c in CASES])
def test_generated_function_lowers_to_the_same_ir(fixture, make, which, kw):
    cfg = make()
    s = cfg.pde_system
    vi = npde.get_vars(s.ivs, s.dvs)
    eq = getattr(s, which[0])[which[1]]
    eq_params = [str(p) for p in s.ps]
    estim = kw.get("param_estim", False)
    ours = npde.lower_equation(eq, vi, {p: i for i, p in enumerate(eq_params)} if estim else {},
                               {} if estim else {p: v for p, v in zip(eq_params, kw.get("default_p", []))}, hoist=False)
    jl = lower_loss_function(parse(_text(fixture)), list(vi.depvars), eq_params, estim, kw.get("default_p"))
    # same taps (network, order, directions) ...
    ours_taps = {(t.net, t.order, tuple(sorted(t.dirs))) for t in ours.taps}
    assert set(jl.taps) == ours_taps
    # ... same rows feeding every tapped network ...
    for k in {t[0] for t in jl.taps}:
        assert jl.net_rows[k] == list(ours.net_rows[k])
    assert jl.dim == len(ours.indvars) and not jl.const_rows
    # ... and the same residual at random points / tap values / parameter values
    rng = np.random.default_rng(5)
    X = rng.uniform(0.1, 0.9, size=(jl.dim, 7))
    params = rng.uniform(0.5, 1.5, size=max(1, len(eq_params)))
    tapval = {key: rng.standard_normal(7) for key in ours_taps}
    r_jl = _run(jl.prog, X, [tapval[k] for k in jl.taps], params)
    r_ours = _run(ours.prog, X, [tapval[(t.net, t.order, tuple(sorted(t.dirs)))] for t in ours.taps], params)
    np.testing.assert_allclose(r_jl, r_ours, rtol=1e-13, atol=1e-14)
    if estim:            # the parameter is a live input of the program, not a folded constant
        r2 = _run(jl.prog, X, [tapval[k] for k in jl.taps], params + 1.0)
        assert not np.allclose(r2, r_jl)


def test_constant_boundary_coordinate_becomes_an_appended_row():
    """QuadratureTraining boundary terms bind the fixed coordinate with fill(value, ...) (get_indvars_ex,
    src/symbolic_utilities.jl:372-386): the network still takes 3 inputs, so the shim appends a constant row."""
    jl = lower_loss_function(parse(_text("cfg4_lid_bc")), ["u", "v", "w", "p"])
    assert jl.taps == [(0, 0, ())] and jl.net_rows[0] == [0, 1, 2] and jl.const_rows == {2: 1.0} and jl.dim == 3
    X = np.random.default_rng(0).uniform(size=(3, 5))
    np.testing.assert_allclose(_run(jl.prog, X, [np.arange(5.0)], [0.0]), np.arange(5.0) - 1.0)


def test_lowering_refuses_what_the_engine_cannot_evaluate():
    """Loud failures, as the Julia pass raises ArgumentError: a function without an opcode, an unknown binding, a
    derivative whose epsilon vectors are not one-hot."""
    base = _text("cfg2_bc")
    bad_fn = base.replace("u(cord1, θ, phi) .- 0.0", "besselj.(u(cord1, θ, phi)) .- 0.0")
    with pytest.raises(ValueError, match="outside the grammar"):
        lower_loss_function(parse(bad_fn), ["u"])
    bad_bind = base.replace("(cord[[1], :], cord[[2], :])", "(cord[[1], :], somewhere(y))")
    with pytest.raises(ValueError, match="unrecognised binding"):
        lower_loss_function(parse(bad_bind), ["u"])
    pde = _text("cfg2_pde").replace("[[0.0001220703125, 0.0], [0.0001220703125, 0.0]]", "[[0.0001220703125, 0.0001220703125], [0.0001220703125, 0.0]]")
    with pytest.raises(AssertionError):
        lower_loss_function(parse(pde), ["u"])


def test_common_subexpressions_are_emitted_once():
    """The IR is SSA with value numbering: u(cord1, θ, phi) appearing twice is one tap and one instruction."""
    txt = _text("cfg2_bc").replace("u(cord1, θ, phi) .- 0.0", "(*).(u(cord1, θ, phi), u(cord1, θ, phi)) .- sin.(u(cord1, θ, phi))")
    jl = lower_loss_function(parse(txt), ["u"])
    assert jl.taps == [(0, 0, ())] and [i[0] for i in jl.prog].count("tap") == 1
    assert [i[0] for i in jl.prog] == ["tap", "mul", "sin", "sub"]
