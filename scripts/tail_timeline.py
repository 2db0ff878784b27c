"""Where the time of the fused kernel's tail goes (tail.cuh), per rank: globaltimer marks of every CTA at tail entry, after
the grid barrier, after the slice reduction (+ push to the peers) and after the peers' slices were added.  Needs the
instrumented build (python neuralpde.jl_b200/build.py --debug).  Single GPU: python scripts/tail_timeline.py
two GPUs: python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 scripts/tail_timeline.py"""
import os
os.environ.setdefault("PINN_B200_LIB", os.path.join("neuralpde.jl_b200", "lib", "libpinn_b200_debug.so"))
import ctypes as C, sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import torch.distributed as dist
import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
from neuralpde_jl_b200.strategies import GridTraining
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(lr)
if world > 1:
    dist.init_process_group("gloo")
cfg = configs.config2()
if world > 1:
    cfg.strategy = GridTraining([1.0 / 127, 1.0 / (128 * world - 1)])
rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float32, mode="tc_split", device=lr), rank=rank, world=world)
eng = rep.engine
if world > 1:
    uid = [npde.Engine.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    eng.comm_init(uid[0], rank, world)
lib = eng.lib
lib.pinn_debug_tail_marks.argtypes = [C.c_void_p, C.c_void_p]
lib.pinn_debug_tail_marks.restype = C.c_int
assert lib.pinn_debug_tail_marks(eng._h, None) == 0
dev = torch.device("cuda", lr)
th = torch.from_numpy(rep.flat_init_params).to(dev)
g = torch.empty_like(th); terms = torch.empty(eng.n_terms, device=dev); tot = torch.empty(1, device=dev)
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
st = torch.cuda.current_stream().cuda_stream
rows = []
for it in range(12):
    flush.zero_()
    eng.loss_grad_device(th, g, terms, tot, None, st)
    buf = np.zeros(256 * 4, dtype=np.int64)
    assert lib.pinn_debug_tail_marks(eng._h, buf.ctypes.data) == 0
    m = buf.reshape(256, 4)
    m = m[m[:, 0] > 0]
    t0 = m[:, 0].min()
    rows.append([m.shape[0], (m[:, 0].max() - t0) / 1e3, (m[:, 1].max() - m[:, 0].max()) / 1e3, (m[:, 2].max() - m[:, 1].max()) / 1e3,
                 (m[:, 3].max() - m[:, 2].max()) / 1e3, (m[:, 3].max() - m[:, 0].max()) / 1e3])
r = np.median(np.array(rows[4:]), axis=0)
print("rank %d/%d: CTAs %d | entry spread (first..last CTA done with its tiles) %.1f us | barrier after last entry %.1f us | "
      "slice reduce (+push) %.1f us | wait for peers + add %.1f us | tail after last entry %.1f us" % (rank, world, r[0], r[1], r[2], r[3], r[4], r[5]), flush=True)
if world > 1:
    dist.barrier()
