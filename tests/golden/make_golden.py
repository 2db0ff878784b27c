"""Generate the golden vectors under tests/golden/ from the float64 oracle (exact-tap mode; the
finite-difference reference mode is stored alongside as a cross-check).

    python tests/golden/make_golden.py

The reference (Julia) cannot run in this image and ships no loss-value goldens for this path
(SURVEY section 4), so these are produced by the oracle restatement and pinned by the
reference's own known-answer tests (tests/test_oracle_pinning.py)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import CASES, point_sets          # noqa: E402
from helpers import oracle_eval              # noqa: E402


def main():
    only = sys.argv[1:]          # optional: regenerate only the named cases
    for name, make in CASES.items():
        if only and name not in only:
            continue
        cfg = make()
        theta = cfg.init_params(np.float64, seed=1)
        sets, qw, qs = point_sets(cfg)
        quad = None if qw is None else (qw, qs)
        L, T, G = oracle_eval(cfg, theta, "exact", sets, quad)
        Lfd, Tfd, Gfd = oracle_eval(cfg, theta, "fd", sets, quad)
        out = {"theta": theta, "total": L, "terms": T, "grad": G, "total_fd": Lfd, "terms_fd": Tfd, "grad_fd": Gfd,
               "n_sets": len(sets)}
        for i, s in enumerate(sets):
            out["set_%d" % i] = s
        if qw is not None:
            for i, w in enumerate(qw):
                out["qw_%d" % i] = w
            out["qscale"] = np.array(qs)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), **out)
        print("%-12s total %.12g  fd-exact rel %.2e  grad rel %.2e  n_theta %d" % (
            name, L, abs(L - Lfd) / abs(L), np.linalg.norm(G - Gfd) / np.linalg.norm(G), theta.size))


if __name__ == "__main__":
    main()
