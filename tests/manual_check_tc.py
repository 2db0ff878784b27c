"""Verbose tcgen05-path check against the float64 oracle (development aid)."""
import sys, time
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
from neuralpde_jl_b200 import configs
from helpers import engine_eval, oracle_eval, rel

cases = [("cfg1", {}), ("cfg2", dict(n=24, width=16, hidden=2)), ("cfg2", dict(n=24, width=32, hidden=3)),
         ("cfg2", dict(n=30, width=48, hidden=3)), ("cfg2", dict(n=40)), ("cfg2", {})]
which = sys.argv[1:] or ["tc_split", "tc_bf16"]
for name, kw in cases:
    cfg = configs.ALL[name](**kw)
    L = T = G = None
    for mode in which:
        t = time.time()
        rep, total, terms, grad = engine_eval(cfg, np.float32, mode=mode)
        te = time.time() - t
        if L is None:
            L, T, G = oracle_eval(cfg, rep.flat_init_params.astype(np.float64))
        # per-layer gradient errors
        print(name, kw, mode, "loss", total, "oracle", L, "rel %.3e" % (abs(total - L) / abs(L)),
              "terms rel %.3e" % np.max(np.abs(terms - T) / np.maximum(np.abs(T), 1e-30)), "grad rel %.3e" % rel(grad, G),
              "t %.3f" % te, flush=True)
        net = cfg.chains[0]
        o = 0
        for li, l in enumerate(net.layers):
            nw = l.in_dims * l.out_dims
            print("    layer %d W rel %.3e b rel %.3e" % (li, rel(grad[o:o + nw], G[o:o + nw]),
                                                          rel(grad[o + nw:o + nw + l.out_dims], G[o + nw:o + nw + l.out_dims])))
            o += nw + l.out_dims
