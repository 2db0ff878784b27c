"""Extract the reference's embedded Burgers solution table (test/DGM/dgm__burger_s_equation.jl:9-25: MethodOfLines
solution of u_t + u u_x - 0.05 u_xx = 0, u(0,x) = -sin(pi x), u(t,+-1) = 0, on an 11 x 21 (t, x) lattice) into
tests/golden/burgers_ref_table.npz.  Run in the build container, where /root/reference exists:
    python tests/golden/make_burgers_table.py"""
import os
import re

import numpy as np

SRC = "/root/reference/test/DGM/dgm__burger_s_equation.jl"
HERE = os.path.dirname(os.path.abspath(__file__))


def _vector(text, name):
    m = re.search(r"const %s = \[(.*?)\]" % name, text, re.S)
    return m.group(1)


def main():
    text = open(SRC).read()
    ts = np.array([float(v) for v in _vector(text, "BURGER_REF_TS").split(",")])
    xs = np.array([float(v) for v in _vector(text, "BURGER_REF_XS").split(",")])
    rows = [r.strip() for r in _vector(text, "BURGER_REF_U").strip().split(";")]
    u = np.array([[float(v) for v in r.split()] for r in rows if r])
    assert u.shape == (ts.size, xs.size) == (11, 21)
    np.savez(os.path.join(HERE, "burgers_ref_table.npz"), ts=ts, xs=xs, u=u, source=SRC + ":9-25")
    print("wrote burgers_ref_table.npz", u.shape)


if __name__ == "__main__":
    main()
