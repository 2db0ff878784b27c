"""Golden loss / term losses / gradient at the FULL shapes BASELINE.json names (the kernels' tile schedulers, multi-pass
channel splitting and 128 / 256-wide layers at scale), from the float64 oracle in exact-tap mode.

    python tests/golden/make_golden_full.py [case ...]

theta and the point sets are NOT stored: both are regenerated from fixed seeds by tests/cases.py (`FULL_CASES`,
`point_sets`) on the test side; stored are total, terms (float64) and the gradient (float32: the modes tested at these
shapes are judged at 1e-5 .. 1e-2).  A sha256 of theta and of the first point set guards the regeneration."""
import hashlib
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import FULL_CASES, point_sets          # noqa: E402
from helpers import oracle_eval                   # noqa: E402


def digest(a) -> str:
    return hashlib.sha256(np.ascontiguousarray(a, dtype=np.float64).tobytes()).hexdigest()


def main():
    only = sys.argv[1:]
    for name, make in FULL_CASES.items():
        if only and name not in only:
            continue
        t0 = time.time()
        cfg = make()
        theta = cfg.init_params(np.float64, seed=1)
        sets, qw, qs = point_sets(cfg)
        quad = None if qw is None else (qw, qs)
        L, T, G = oracle_eval(cfg, theta, "exact", sets, quad)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), total=L, terms=T, grad=G.astype(np.float32),
                            theta_sha=digest(theta), set0_sha=digest(sets[0]), n_theta=theta.size,
                            n_points=np.array([s.shape[1] for s in sets]))
        print("%-14s total %.12g  n_theta %d  points %s  (%.1f s)" % (name, L, theta.size, [s.shape[1] for s in sets][:4],
                                                                        time.time() - t0), flush=True)


if __name__ == "__main__":
    main()
