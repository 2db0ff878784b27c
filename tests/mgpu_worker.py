"""Worker for tests/test_gpu_multi.py: each rank evaluates its shard on its own GPU; the engine all-reduces
[grad | term losses] over NCCL.  Launched with torch.distributed.run."""
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
if rank == 0:
    np.savez(out, tot=tot, terms=terms, g=g, tot2=tot2, terms2=terms2)
dist.barrier()
dist.destroy_process_group()
