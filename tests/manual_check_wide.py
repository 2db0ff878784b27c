"""First-light check of the 128-wide tensor path against the committed goldens, with a per-parameter-block error
breakdown (localises a wrong layer / phase in one run)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import neuralpde_jl_b200 as npde
from cases import CASES
from helpers import engine_eval_sets, load_golden, rel

names = sys.argv[1:] or ["poisson1d_wide", "burgers_wide", "cfg5_wide"]
for name in names:
    g, sets, qw = load_golden(name)
    cfg = CASES[name]()
    for want_grad in (False, True):
        try:
            rep, total, terms, grad = engine_eval_sets(cfg, np.float32, sets, qw, mode="tc_bf16", theta=g["theta"], want_grad=want_grad)
        except Exception as ex:
            print(name, "FAILED:", repr(ex)[:300]); break
        print("%s grad=%s total %.8g (golden %.8g, rel %.2e)" % (name, want_grad, total, float(g["total"]),
                                                                 abs(total - float(g["total"])) / abs(float(g["total"]))))
        print("   terms", np.array2string(terms, precision=5), " golden", np.array2string(np.asarray(g["terms"]), precision=5))
        if want_grad:
            print("   grad rel", rel(grad, g["grad"]))
            off = 0
            dims = cfg.chains[0].dims
            for l in range(len(dims) - 1):
                nw = dims[l] * dims[l + 1]
                print("   layer %d  W rel %.3e   b rel %.3e   |W_gold| %.3e" % (
                    l, rel(grad[off:off + nw], g["grad"][off:off + nw]), rel(grad[off + nw:off + nw + dims[l + 1]], g["grad"][off + nw:off + nw + dims[l + 1]]),
                    float(np.linalg.norm(g["grad"][off:off + nw]))))
                off += nw + dims[l + 1]
        rep.engine.close()
