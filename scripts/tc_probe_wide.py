"""Pin the descriptor conventions the wide (128-column) tensor path relies on: K-major B with N = 128 rows,
MN-major operands whose M / N extent of 128 spans two 64-column tiles (LBO = tile stride)."""
import sys
import numpy as np
sys.path.insert(0, "."); sys.path.insert(0, "scripts")
from tc_probe import probe, swz_image, idesc, bf16_round, err   # noqa: E402 (runs the base probe too)

rng = np.random.default_rng(1)
TB = 16384
A = bf16_round(rng.standard_normal((128, 64)))
print("== (W1) K-major A[128x64] x K-major B[128(n) x 64]: D[128x128] = A @ B.T")
B = bf16_round(rng.standard_normal((128, 64)))
D = probe(swz_image(A), swz_image(B), 0, 0, 0, 1024, 0, 1024, 32, 32, 4, idesc(128, 128, 0, 0), 128)
print("  err", err(D, A @ B.T))
print("== (W2) dgrad N=128: A K-major [128 x 64(o)], B MN-major from two tiles W[:, 0:64], W[:, 64:128] (rows = o)")
W = bf16_round(rng.standard_normal((128, 128)))          # rows o, cols n
img = np.concatenate([swz_image(W[:, :64]), swz_image(W[:, 64:])])
for lbo in (TB, 0, 1024):
    D = probe(swz_image(A), img, 0, 0, 0, 1024, lbo, 1024, 32, 2048, 4, idesc(128, 128, 0, 1), 128)
    print("  b_lbo", lbo, "err cols0-63", err(D[:, :64], A @ W[:64, :64]), "cols64-127", err(D[:, 64:], A @ W[:64, 64:]))
print("== (W3) wgrad 128x128: A MN-major Z[128(p) x 128(m)] two tiles, B MN-major H[128(p) x 128(n)] two tiles: D = Z.T @ H")
Z = bf16_round(rng.standard_normal((128, 128)))
H = bf16_round(rng.standard_normal((128, 128)))
zi = np.concatenate([swz_image(Z[:, :64]), swz_image(Z[:, 64:])])
hi = np.concatenate([swz_image(H[:, :64]), swz_image(H[:, 64:])])
ref = Z.T @ H
for (al, bl) in ((TB, TB), (0, TB), (TB, 0)):
    D = probe(zi, hi, 0, 0, al, 1024, bl, 1024, 2048, 2048, 8, idesc(128, 128, 1, 1), 128)
    print("  a_lbo", al, "b_lbo", bl, "err", err(D, ref), " q00 %.3g q01 %.3g q10 %.3g q11 %.3g" % (
        err(D[:64, :64], ref[:64, :64]), err(D[:64, 64:], ref[:64, 64:]), err(D[64:, :64], ref[64:, :64]), err(D[64:, 64:], ref[64:, 64:])))
print("== (W4) layer-0 wgrad shape: A MN-major Z two tiles (M=128), B MN-major [128(p) x 16]: D[128 x 16]")
X = bf16_round(rng.standard_normal((128, 64)))
D = probe(zi, swz_image(X), 0, 0, TB, 1024, 0, 1024, 2048, 2048, 8, idesc(128, 16, 1, 1), 16)
print("  err", err(D, Z.T @ X[:, :16]))
print("== (W5) column sums with a 1 KB constant B atom: B MN-major, SBO = 0, k-step 0, N = 16: D[o][0] = sum_p Z[p][o]")
ones = np.zeros((8, 64), dtype=np.float32); ones[:, 0] = 1.0
for (sbo, step) in ((0, 0), (1024, 0)):
    try:
        D = probe(zi, swz_image(ones), 0, 0, TB, 1024, 0, sbo, 2048, step, 8, idesc(128, 16, 1, 1), 16)
        print("  b_sbo", sbo, "b_step", step, "err col0", err(D[:, 0], Z.sum(axis=0)), "max |cols 1..15|", float(np.abs(D[:, 1:]).max()))
    except Exception as e:
        print("  fail", sbo, step, e)
