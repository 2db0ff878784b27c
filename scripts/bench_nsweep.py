"""N-sweep of the device-resident loss+grad step: 2-D Poisson, 4x64 tanh MLP, n x n grids."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "tests")
import torch
import neuralpde_jl_b200 as npde
from neuralpde_jl_b200 import configs

dev = torch.device("cuda")
print("| grid | PDE points | mode | ms / step | M pts/s | TFLOP/s (algorithmic) |")
print("|---|---|---|---|---|---|")
for n in (128, 256, 512, 1024, 2048):
    for mode in (sys.argv[1:] or ("tc_split", "tc_bf16", "ffma")):
        if mode == "ffma" and n > 1024: continue
        cfg = configs.config2(n=n)
        rep = npde.symbolic_discretize(cfg.pde_system, cfg.discretization(dtype=np.float32, mode=mode))
        eng = rep.engine
        th = torch.from_numpy(rep.flat_init_params).to(dev)
        g = torch.empty_like(th); terms = torch.empty(eng.n_terms, device=dev); tot = torch.empty(1, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        for _ in range(3):
            eng.loss_grad_device(th, g, terms, tot, None, st)
        torch.cuda.synchronize()
        it = 20 if n <= 512 else 5
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(it):
            eng.loss_grad_device(th, g, terms, tot, None, st)
        t1.record(); torch.cuda.synchronize()
        ms = t0.elapsed_time(t1) / it
        print("| %dx%d | %d | %s | %.4f | %.1f | %.1f |" % (n, n, n * n, mode, ms, n * n / ms / 1e3, eng.flops_per_eval() / ms / 1e9), flush=True)
        eng.close()
