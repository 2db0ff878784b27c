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
@pytest.mark.parametrize("mode,tol,path", [("ffma", 1e-11, "p2p"), ("tc_split", 2e-6, "p2p"), ("tc_split", 2e-6, "nccl"),
                                           ("tc_bf16", 2e-6, "p2p")])
def test_two_rank_allreduce_matches_single_rank(tmp_path, mode, tol, path):
    """path p2p: the sum over ranks runs in the fused kernel's tail over peer memory (asserted via pinn_comm_info);
    path nccl: PINN_B200_NO_P2P=1 forces fused kernel -> ncclAllReduce -> unpack."""
    out = str(tmp_path / "r0.npz")
    world = min(torch.cuda.device_count(), 8 if path == "p2p" and mode == "tc_split" else 2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1",
           "--master-port", "29561", os.path.join(ROOT, "tests", "mgpu_worker.py"), mode, out]
    env = dict(os.environ)
    if path == "nccl":
        env["PINN_B200_NO_P2P"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    res = np.load(out)
    assert int(res["fused"]) == (1 if path == "p2p" else 0), str(res["why"])
    dtype = np.float64 if mode == "ffma" else np.float32
    cfg = configs.config2(n=48, width=32, hidden=3)
    rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode))
    tot, terms, g = rep.engine.loss_grad_host(rep.flat_init_params, None, True)
    assert abs(float(res["tot"]) - tot) <= tol * abs(tot)
    np.testing.assert_allclose(res["terms"], terms, rtol=10 * tol)
    assert rel(res["g"], g) < 10 * tol
    assert abs(float(res["tot2"]) - tot) <= tol * abs(tot)
    assert abs(float(res["tot3"]) - tot) <= tol * abs(tot) and rel(res["g3"], g) < 10 * tol
    if path == "p2p":
        # the replicated device-resident Adam loop equals the single-GPU loop (same total gradient up to summation order)
        rep.engine.adam_begin(rep.flat_init_params, 1e-3)
        rep.engine.adam_iterate(6)
        assert rel(res["th_adam"], rep.engine.adam_theta()) < (1e-9 if mode == "ffma" else 2e-4)
