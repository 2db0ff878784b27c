"""CPU oracle (test infrastructure only -- see oracle/reference.py).  Never imported by the
product package ``neuralpde.jl_b200``."""
