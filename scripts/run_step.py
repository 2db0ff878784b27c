"""Run a few device-resident loss+grad steps of one config / mode (profiling target for ncu)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
mode = sys.argv[2] if len(sys.argv) > 2 else "tc_bf16"
n = int(sys.argv[3]) if len(sys.argv) > 3 else 4
cfg = getattr(configs, "config" + which[-1])()
rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float32, mode=mode))
eng = rep.engine
dev = torch.device("cuda")
th = torch.from_numpy(rep.flat_init_params).to(dev)
g = torch.empty_like(th); terms = torch.empty(eng.n_terms, device=dev); tot = torch.empty(1, device=dev)
if hasattr(rep.strategy, "points"):
    rep.loss_functions.full_loss_function(rep.flat_init_params)
for _ in range(n):
    eng.loss_grad_device(th, g, terms, tot, None, torch.cuda.current_stream().cuda_stream)
torch.cuda.synchronize()
print("loss", float(tot.item()))
