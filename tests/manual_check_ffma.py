"""Verbose GPU-vs-oracle check used while developing (not part of the test suite)."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from neuralpde_jl_b200 import configs
from helpers import engine_eval, oracle_eval, rel

cases = [("cfg1", {}, np.float64), ("cfg2", dict(n=24, width=16, hidden=2), np.float64),
         ("cfg2", dict(n=24, width=16, hidden=2), np.float32), ("cfg2", {}, np.float64), ("cfg2", {}, np.float32)]
for name, kw, dt in cases:
    cfg = configs.ALL[name](**kw)
    t = time.time()
    rep, total, terms, grad = engine_eval(cfg, dt)
    te = time.time() - t
    L, T, G = oracle_eval(cfg, rep.flat_init_params.astype(np.float64))
    print(name, kw, np.dtype(dt).name, "engine", total, "oracle", L, "rel", abs(total - L) / abs(L),
          "terms rel", np.max(np.abs(terms - T) / np.abs(T)), "grad rel", rel(grad, G), "t_engine %.3f" % te, flush=True)
