"""Worker for tests/test_gpu_multi.py: each rank evaluates its shard on its own GPU; the gradient and the term losses are
summed over the ranks inside the fused kernel over peer memory (or by ncclAllReduce with PINN_B200_NO_P2P=1), then a
few steps of the device-resident Adam loop run on every rank.  Launched with torch.distributed.run."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import neuralpde_jl_b200 as npde          # noqa: E402
from neuralpde_jl_b200 import configs     # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dist.init_process_group("gloo")
mode, out = sys.argv[1], sys.argv[2]
dtype = np.float64 if mode == "ffma" else np.float32
cfg = configs.config2(n=48, width=32, hidden=3)
rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=dtype, mode=mode, device=int(os.environ["LOCAL_RANK"])),
                               rank=rank, world=world)
uid = [npde.Engine.comm_unique_id() if rank == 0 else None]
dist.broadcast_object_list(uid, src=0)
rep.engine.comm_init(uid[0], rank, world)
tot, terms, g = rep.engine.loss_grad_host(rep.flat_init_params, None, True)
tot2, terms2, _ = rep.engine.loss_grad_host(rep.flat_init_params, None, False)      # loss-only path of the allreduce
fused, why = rep.engine.comm_info()
# device-resident Adam over the ranks (peer-memory path only): every rank must end with the same theta bit for bit
th_adam, adam_loss = np.zeros(0), float("nan")
if fused:
    rep.engine.adam_begin(rep.flat_init_params, 1e-3)
    rep.engine.adam_iterate(3)
    adam_loss, _ = rep.engine.adam_iterate(3)           # second call replays the captured graph
    th_adam = rep.engine.adam_theta()
    gathered = [None] * world
    dist.all_gather_object(gathered, th_adam.tobytes())
    assert all(b == gathered[0] for b in gathered), "replicas diverged"
# repeated evaluations: the flag / parity protocol over many steps
for _ in range(20):
    tot3, _, g3 = rep.engine.loss_grad_host(rep.flat_init_params, None, True)
if rank == 0:
    np.savez(out, tot=tot, terms=terms, g=g, tot2=tot2, terms2=terms2, fused=int(fused), why=why, th_adam=th_adam,
             adam_loss=adam_loss, tot3=tot3, g3=g3)
dist.barrier()
dist.destroy_process_group()
