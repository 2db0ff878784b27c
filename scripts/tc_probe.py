"""Sweep tcgen05 descriptor hypotheses on the GPU (uses pinn_debug_mma_probe)."""
import os
os.environ.setdefault("PINN_B200_LIB", os.path.join("neuralpde.jl_b200", "lib", "libpinn_b200_debug.so"))   # build.py --debug
import ctypes as C, itertools, sys
import numpy as np
sys.path.insert(0, ".")
import torch
import neuralpde_jl_b200 as npde
lib = npde.load_library()
lib.pinn_debug_mma_probe.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p]
lib.pinn_debug_mma_probe.restype = C.c_int

def bf16_round(x):
    return torch.tensor(x, dtype=torch.float32).to(torch.bfloat16).to(torch.float32).numpy()

def bf16_bits(x):
    return torch.tensor(x, dtype=torch.float32).to(torch.bfloat16).view(torch.int16).numpy().astype(np.uint16)

def swz_image(T):
    """T: (rows, 64) float -> bytes image of a 128B-swizzled bf16 tile."""
    rows = T.shape[0]
    bits = bf16_bits(T)
    img = np.zeros(rows * 64, dtype=np.uint16)
    r = np.arange(rows)[:, None]; c = np.arange(64)[None, :]
    off = r * 128 + ((((c >> 3) ^ (r & 7)) & 7) << 4) + ((c & 7) << 1)
    img[(off // 2).ravel()] = bits.ravel()
    return img.view(np.uint8)

def idesc(m, n, a_mn, b_mn):
    return (1 << 4) | (1 << 7) | (1 << 10) | (a_mn << 15) | (b_mn << 16) | ((n >> 3) << 17) | ((m >> 4) << 24)

def probe(a_img, b_img, a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, nk, idsc, ncols):
    p = np.array([a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, nk, idsc, ncols], dtype=np.uint32)
    out = np.zeros((128, ncols), dtype=np.float32)
    a = np.ascontiguousarray(a_img); b = np.ascontiguousarray(b_img)
    rc = lib.pinn_debug_mma_probe(a.ctypes.data, a.nbytes, b.ctypes.data, b.nbytes, p.ctypes.data, out.ctypes.data)
    if rc: raise RuntimeError("probe rc=%d" % rc)
    return out

rng = np.random.default_rng(0)
A = bf16_round(rng.standard_normal((128, 64)))     # rows x 64
B = bf16_round(rng.standard_normal((128, 64)))
def err(x, y): return float(np.max(np.abs(x - y)))

print("== (1) K-major A[128x64] x K-major B[64x64]: D = A @ B[:64].T")
for lbo in (0, 16, 1024):
    D = probe(swz_image(A), swz_image(B), 0, 0, lbo, 1024, lbo, 1024, 32, 32, 4, idesc(128, 64, 0, 0), 64)
    print("  lbo", lbo, "err", err(D, A @ B[:64].T))
print("== (2) N=32, B rows 32..63 (b_off 4096)")
D = probe(swz_image(A), swz_image(B), 0, 4096, 0, 1024, 0, 1024, 32, 32, 4, idesc(128, 32, 0, 0), 32)
print("  err", err(D, A @ B[32:64].T))
print("== (3) dgrad: A K-major [128 x 64(o)], B MN-major from W[64(o) rows x 64(n) cols]: D = A @ W")
W = B[:64]
for (lbo, sbo, step) in itertools.product((0, 1024, 8192), (1024, 128), (2048, 256, 32)):
    try:
        D = probe(swz_image(A), swz_image(W), 0, 0, 0, 1024, lbo, sbo, 32, step, 4, idesc(128, 64, 0, 1), 64)
        print("  b_lbo", lbo, "b_sbo", sbo, "b_step", step, "err", err(D, A @ W))
    except Exception as e:
        print("  fail", lbo, sbo, step, e)
print("== (4) wgrad: A MN-major Z[128(p) x 64(m)], B MN-major H[128(p) x 64(n)]: D = Z.T @ H   (M=128 dup via a_lbo)")
Z, H = A, B
ref = Z.T @ H
for a_lbo in (0, 1024):
    D = probe(swz_image(Z), swz_image(H), 0, 0, a_lbo, 1024, 0, 1024, 2048, 2048, 8, idesc(128, 64, 1, 1), 64)
    print("  a_lbo", a_lbo, "rows0-63 err", err(D[:64], ref), "rows64-127 vs dup err", err(D[64:], ref))
print("== (5) wgrad with true M=64: where do rows land?")
D = probe(swz_image(Z), swz_image(H), 0, 0, 0, 1024, 0, 1024, 2048, 2048, 8, idesc(64, 64, 1, 1), 64)
for lane0 in range(0, 128, 16):
    blk = D[lane0:lane0 + 16]
    best = min(range(0, 64, 16), key=lambda r: err(blk, ref[r:r + 16]))
    print("  lanes %3d-%3d ~ rows %2d.. err %.3g" % (lane0, lane0 + 15, best, err(blk, ref[best:best + 16])))
print("== (6) split check: A two k-blocks? K=16 only (1 kstep) offsets")
D = probe(swz_image(A), swz_image(B), 0, 0, 0, 1024, 0, 1024, 32, 32, 1, idesc(128, 64, 0, 0), 64)
print("  1 kstep err", err(D, A[:, :16] @ B[:64, :16].T))
