"""Two GPUs, one process each: sharded point sets + the engine's NCCL allreduce reproduce the single-GPU loss and
gradient.  Skipped on a one-GPU box (the CPU twin is tests/test_distributed_gloo.py)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from helpers import rel

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("mode,tol", [("ffma", 1e-11), ("tc_split", 2e-6)])
def test_two_rank_allreduce_matches_single_rank(tmp_path, mode, tol):
    out = str(tmp_path / "r0.npz")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tests", "mgpu_worker.py"), mode, out]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    res = np.load(out)
    dtype = np.float64 if mode == "ffma" else np.float32
    cfg = configs.config2(n=48, width=32, hidden=3)
    rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode))
    tot, terms, g = rep.engine.loss_grad_host(rep.flat_init_params, None, True)
    assert abs(float(res["tot"]) - tot) <= tol * abs(tot)
    np.testing.assert_allclose(res["terms"], terms, rtol=10 * tol)
    assert rel(res["g"], g) < 10 * tol
    assert abs(float(res["tot2"]) - tot) <= tol * abs(tot)
