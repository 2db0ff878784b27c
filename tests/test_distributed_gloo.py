"""N > 1 host logic on CPU: world_size-2 gloo.  Each rank evaluates its contiguous shard of every
term's point set (oracle as the compute stand-in), the ranks all-reduce ONE packed buffer
[grad | per-term sum r^2] and recover the unsharded loss and gradient."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from neuralpde_jl_b200.strategies import shard_range
from oracle import reference as R
from cases import CASES
from helpers import load_golden


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, name, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    g, sets, _ = load_golden(name)
    cfg = CASES[name]()
    sys_ = cfg.pde_system
    prob = R.Problem(sys_, cfg.chain_specs(), derivative="exact")
    theta = torch.tensor(g["theta"], requires_grad=True)
    eqs = list(sys_.eqs) + list(sys_.bcs)
    n_terms = len(eqs)
    # local weighted objective: sum_k (1/N_k) sum_{i in shard} r_i^2  (N_k = global count)
    sums, obj = [], 0.0
    for k, (eq, s) in enumerate(zip(eqs, sets)):
        lo, hi = shard_range(s.shape[1], rank, world)
        if hi > lo:
            r = prob.residual(eq, torch.as_tensor(s[:, lo:hi]), theta)
            sk = torch.sum(r * r)
        else:
            sk = torch.zeros(())
        sums.append(sk / s.shape[1])
        obj = obj + sk / s.shape[1]
    (grad,) = torch.autograd.grad(obj, theta)
    packed = torch.cat([grad, torch.stack([t.detach() for t in sums])])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)          # the one collective per step
    if rank == 0:
        np.savez(out, grad=packed[:-n_terms].numpy(), terms=packed[-n_terms:].numpy())
    dist.destroy_process_group()


def test_sharded_allreduce_recovers_unsharded(tmp_path):
    name = "cfg2_small"
    out = str(tmp_path / "r0.npz")
    mp.spawn(_worker, args=(2, _free_port(), name, out), nprocs=2, join=True)
    res = np.load(out)
    g, _, _ = load_golden(name)
    np.testing.assert_allclose(res["terms"], g["terms"], rtol=1e-12)
    np.testing.assert_allclose(res["grad"], g["grad"], rtol=1e-9, atol=1e-13)
    assert abs(res["terms"].sum() - float(g["total"])) < 1e-12 * float(g["total"])
