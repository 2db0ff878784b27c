// tc_kernel.cu -- fused PINN loss+gradient kernel, tcgen05 tensor-core path (sm_100a).
//
// One CTA (512 threads) owns a tile of 128 collocation points; a point is a TMEM lane and a
// row of every operand tile.  The hidden->hidden Dense layers run on the 5th-generation
// tensor cores (tcgen05.mma, bf16 operands from 128B-swizzled shared-memory tiles, fp32
// accumulators in TMEM); every derivative channel (value, d/dx_i, d2/dx_i dx_j) is its own
// 128-row M block that shares the same weight operand.  The epilogue (bias + activation +
// forward-mode tap chain rule, or its reverse) runs on the CUDA cores straight out of TMEM
// and re-packs the result as the next GEMM's bf16 operand tile.
//
//   forward, per tensor layer l :  D_c[128 x n_out] = H_c[128 x n_in] * W_l^T        (A, B K-major)
//   backward, per tensor layer l:  Z_c   (recompute, 32-column groups)  = H_c * W_l^T
//                                  Hbar_c[128 x n_in] = Zbar_c[128 x n_out] * W_l     (B MN-major)
//                                  Wbar_l[n_out x n_in] = sum_c Zbar_c^T * H_c         (A, B MN-major)
//                                  bbar_l[n_out]        = Zbar_0^T * 1                  (B = constant ones atom)
//   last layer                  :  wbar_L[n]            = sum_c H_c^T * ubar_c          (B = (hi, lo) pairs of ubar)
//
// The first (d -> n) and last (n -> 1) layers are tiny and stay on the CUDA cores inside the
// same epilogues.  Arithmetic modes: PINN_MODE_TC_BF16 (one MMA per product) and
// PINN_MODE_TC_SPLIT (forward operands split into bf16 hi + lo, three MMAs per product,
// which restores ~fp32 accuracy of the loss; the reverse sweep uses the hi parts).
//
// Replaces the same reference functions as the FFMA path (see ffma_kernel.cuh).
#include "tc_common.cuh"
#include "tail.cuh"

#ifndef PINN_TC_PREFETCH
#define PINN_TC_PREFETCH 0      // 1: software-pipelined TMEM loads in the tensor-layer epilogues (measurement variant)
#endif

namespace pinn {

// CTA-wide constants kept in shared memory so the per-network passes (separate functions) do not
// drag a context struct through local memory
struct CtaShared {
  uint32_t tmem;
  int split, tl_max, off_P, off_Q, off_misc, off_ones, mx_dim, mx_taps;
  float* partial;
  uint8_t* stash;
  const float* theta;
  long long* dbg;          // optional phase-timestamp buffer (CTA 0, thread 0)
  int dbg_n;
  TcNetSmem nets[PINN_MAX_NETS];
};


// first-layer pre-activations of neuron o (channel vector zz); fpa = shared-memory address of the fp32 block
template <int N1, int N2>
__device__ __forceinline__ void first_layer_elem(uint32_t fpa, const PassInfo<N1, N2>& pi, const float (&x)[PINN_MAX_IN],
                                                 int o, float* zz) {
  float s = lds_f32(fpa + (FP_B1 + o) * 4);
  const uint32_t wa = fpa + (FP_W1 + o * 8) * 4;
  if (pi.d_in <= 3) {            // common 1-D / 2-D / 3-D problems: no predicated tail
    s = fmaf(lds_f32(wa), x[0], s);
    if (pi.d_in >= 2) s = fmaf(lds_f32(wa + 4), x[1], s);
    if (pi.d_in == 3) s = fmaf(lds_f32(wa + 8), x[2], s);
  } else {
#pragma unroll
    for (int k = 0; k < PINN_MAX_IN; ++k)
      if (k < pi.d_in) s = fmaf(lds_f32(wa + k * 4), x[k], s);
  }
  zz[0] = s;
#pragma unroll
  for (int j = 0; j < N1; ++j) zz[1 + j] = lds_f32(fpa + (FP_W1 + o * 8 + pi.dir1[j]) * 4);
#pragma unroll
  for (int j = 0; j < N2; ++j) zz[1 + N1 + j] = 0.f;
}


// ---- granule loops (4 columns x all channels per step), specialised on the activation kind -------------
// (inlined into the per-network passes; a noinline callee only gets the ABI scratch registers and spills)
struct LoopCtx {
  uint32_t fp;            // shared-memory address of the network's fp32 parameter block
  uint32_t bt;            // shared-memory address of the current tensor layer's bias
  uint32_t tP;            // shared-memory address of the hi operand tiles
  uint32_t tQ;            // shared-memory address of the lo operand tiles (forward split)
  float* gb;              // bias gradient of the current layer (CTA partial)
  float* gw;              // weight gradient of the first layer (CTA partial)
  uint32_t taddr;         // tmem base + lane quadrant
  int act, split, p, lane, g0, g1, c0, flag;
};

// layer 0 forward: coordinates -> H^0 tiles (+ last-layer dot when there is no tensor layer: flag)
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void l0_fwd_loop(const LoopCtx lc, const PassInfo<N1, N2> pi, const float* xp, float* up) {
  constexpr int C = 1 + N1 + N2;
  float x[PINN_MAX_IN], u[C];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = xp[k];
#pragma unroll
  for (int c = 0; c < C; ++c) u[c] = up[c];
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float h[C][GW];
#pragma unroll
    for (int i = 0; i < GW; i += 2) {
      float za[C], zb2[C];
      first_layer_elem<N1, N2>(lc.fp, pi, x, g * GW + i, za);
      first_layer_elem<N1, N2>(lc.fp, pi, x, g * GW + i + 1, zb2);
      P2 zz[C], hv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) zz[c] = mk2(za[c], zb2[c]);
      chain_fwd<N1, N2, PURE, AK, P2>(lc.act, pi.ch, zz, hv);
#pragma unroll
      for (int c = 0; c < C; ++c) { h[c][i] = hv[c].v.x; h[c][i + 1] = hv[c].v.y; }
      if (lc.flag) {
        const float w0 = lds_f32(lc.fp + (FP_WL + g * GW + i) * 4), w1 = lds_f32(lc.fp + (FP_WL + g * GW + i + 1) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) u[c] = fmaf(w1, hv[c].v.y, fmaf(w0, hv[c].v.x, u[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(lc.tP + c * kTileBytes, lc.tQ + c * kTileBytes, lc.p, g * GW, h[c], lc.split != 0);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) up[c] = u[c];
}

// tensor layer forward epilogue: TMEM accumulators -> bias + activation chain -> next operand tiles
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tl_fwd_loop(const LoopCtx lc, const Chan<N1, N2> ch, float* up) {
  constexpr int C = 1 + N1 + N2;
  float u[C];
#pragma unroll
  for (int c = 0; c < C; ++c) u[c] = up[c];
#if PINN_TC_PREFETCH
  // software pipeline: the TMEM loads of granule g + 1 are in flight while granule g is evaluated
  float zn[C][GW];
#pragma unroll
  for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + lc.g0 * GW, zn[c]);
#endif
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float z[C][GW];
#if PINN_TC_PREFETCH
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < GW; ++i) z[c][i] = zn[c][i];
    if (g + 1 < lc.g1) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + (g + 1) * GW, zn[c]);
    }
#else
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + g * GW, z[c]);
    tc::tmem_ld_wait();
#endif
#pragma unroll
    for (int i = 0; i < GW; i += 2) {
      P2 zz[C], hv[C];
      zz[0] = mk2(z[0][i] + lds_f32(lc.bt + (g * GW + i) * 4), z[0][i + 1] + lds_f32(lc.bt + (g * GW + i + 1) * 4));
#pragma unroll
      for (int c = 1; c < C; ++c) zz[c] = mk2(z[c][i], z[c][i + 1]);
      chain_fwd<N1, N2, PURE, AK, P2>(lc.act, ch, zz, hv);
#pragma unroll
      for (int c = 0; c < C; ++c) { z[c][i] = hv[c].v.x; z[c][i + 1] = hv[c].v.y; }
      if (lc.flag) {
        const float w0 = lds_f32(lc.fp + (FP_WL + g * GW + i) * 4), w1 = lds_f32(lc.fp + (FP_WL + g * GW + i + 1) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) u[c] = fmaf(w1, hv[c].v.y, fmaf(w0, hv[c].v.x, u[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(lc.tP + c * kTileBytes, lc.tQ + c * kTileBytes, lc.p, g * GW, z[c], lc.split != 0);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) up[c] = u[c];
}

// tensor layer backward epilogue for the column group starting at c0: recomputed Z (TMEM Y) and output
// adjoints (TMEM X, or w_last * ubar for the last hidden layer: flag) -> Zbar tiles (the bias gradient is a
// column sum of Zbar_0, taken by one MMA chain against the constant ones atom in net_backward)
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tl_bwd_loop(const LoopCtx lc, const Chan<N1, N2> ch, const float* ubp) {
  constexpr int C = 1 + N1 + N2;
  float ub[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ub[c] = ubp[c];
#if PINN_TC_PREFETCH
  float zn[C][GWB], hn[C][GWB];
#pragma unroll
  for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_Y + c * 32 + lc.g0 * GWB, zn[c]);
  if (!lc.flag) {
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + lc.c0 + lc.g0 * GWB, hn[c]);
  }
#endif
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    const int ocol = lc.c0 + g * GWB;
    float z[C][GWB], hb[C][GWB];
#if PINN_TC_PREFETCH
    tc::tmem_ld_wait();
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
      for (int i = 0; i < GWB; ++i) { z[c][i] = zn[c][i]; hb[c][i] = hn[c][i]; }
    if (g + 1 < lc.g1) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_Y + c * 32 + (g + 1) * GWB, zn[c]);
      if (!lc.flag) {
#pragma unroll
        for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + ocol + GWB, hn[c]);
      }
    }
#else
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_Y + c * 32 + g * GWB, z[c]);
    if (!lc.flag) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + ocol, hb[c]);
    }
    tc::tmem_ld_wait();
#endif
    if (lc.flag) {
#pragma unroll
      for (int i = 0; i < GWB; ++i) {
        const float wl = lds_f32(lc.fp + (FP_WL + ocol + i) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) hb[c][i] = wl * ub[c];
      }
    }
#pragma unroll
    for (int i = 0; i < GWB; i += 2) {
      P2 zz[C], hv[C], zv[C];
      zz[0] = mk2(z[0][i] + lds_f32(lc.bt + (ocol + i) * 4), z[0][i + 1] + lds_f32(lc.bt + (ocol + i + 1) * 4));
#pragma unroll
      for (int c = 1; c < C; ++c) zz[c] = mk2(z[c][i], z[c][i + 1]);
#pragma unroll
      for (int c = 0; c < C; ++c) hv[c] = mk2(hb[c][i], hb[c][i + 1]);
      chain_bwd<N1, N2, PURE, AK, P2>(lc.act, ch, zz, hv, zv);
#pragma unroll
      for (int c = 0; c < C; ++c) { hb[c][i] = zv[c].v.x; hb[c][i + 1] = zv[c].v.y; }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(lc.tP + c * kTileBytes, lc.tP, lc.p, ocol, hb[c], false);
  }
}

// layer 0 backward, tensor-core variant: adjoints of H^0 (TMEM X) -> Zbar^0 tiles (value + first-derivative
// channels; bf16 hi) in P.  The weight / bias gradient is then one small MMA chain against the augmented
// coordinate tiles (see net_backward), so no cross-lane reductions are needed here.
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void l0_bwd_store_loop(const LoopCtx lc, const PassInfo<N1, N2> pi, const float* xp) {
  constexpr int C = 1 + N1 + N2;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = xp[k];
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float hb[C][GWB];
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + g * GWB, hb[c]);
    tc::tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < GWB; i += 2) {
      float za[C], zb2[C];
      first_layer_elem<N1, N2>(lc.fp, pi, x, g * GWB + i, za);
      first_layer_elem<N1, N2>(lc.fp, pi, x, g * GWB + i + 1, zb2);
      P2 zz[C], hv[C], zv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { zz[c] = mk2(za[c], zb2[c]); hv[c] = mk2(hb[c][i], hb[c][i + 1]); }
      chain_bwd<N1, N2, PURE, AK, P2>(lc.act, pi.ch, zz, hv, zv);
#pragma unroll
      for (int c = 0; c <= N1; ++c) { hb[c][i] = zv[c].v.x; hb[c][i + 1] = zv[c].v.y; }
    }
#pragma unroll
    for (int c = 0; c <= N1; ++c) store_half(lc.tP + c * kTileBytes, lc.tP, lc.p, g * GWB, hb[c], false);
  }
}

// layer 0 backward: adjoints of H^0 (TMEM X, or w_last * ubar: flag) -> first-layer weight / bias gradient
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void l0_bwd_loop(const LoopCtx lc, const PassInfo<N1, N2> pi, const float* xp, const float* ubp) {
  constexpr int C = 1 + N1 + N2;
  float x[PINN_MAX_IN], ub[C];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = xp[k];
#pragma unroll
  for (int c = 0; c < C; ++c) ub[c] = ubp[c];
  const int re = reduceg_elem(lc.lane);
  const bool rlead = reduceg_lead(lc.lane);
  const int n1w = pi.n1w;
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float hb[C][GW];
    if (!lc.flag) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmem_ldg(lc.taddr + TM_X + c * 64 + g * GW, hb[c]);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < GW; ++i) {
        const float wl = lds_f32(lc.fp + (FP_WL + g * GW + i) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) hb[c][i] = wl * ub[c];
      }
    }
    float zv0[GW], zvd[N1 > 0 ? N1 : 1][GW];
#pragma unroll
    for (int i = 0; i < GW; ++i) {
      // scalar here: the packed form raises the register pressure of this loop past the 128-register budget
      float zz[C], hv[C], zv[C];
      first_layer_elem<N1, N2>(lc.fp, pi, x, g * GW + i, zz);
#pragma unroll
      for (int c = 0; c < C; ++c) hv[c] = hb[c][i];
      chain_bwd<N1, N2, PURE, AK, float>(lc.act, pi.ch, zz, hv, zv);
      zv0[i] = zv[0];
#pragma unroll
      for (int jd = 0; jd < N1; ++jd) zvd[jd][i] = zv[1 + jd];
    }
    // Wbar_0[o][k] = sum_p zbar_0 x_k + zbar_(channel of direction k);  bbar_0[o] = sum_p zbar_0
    const float bs = warp_reduceg(zv0, lc.lane);
    if (rlead) atomicAdd(lc.gb + g * GW + re, bs);
#pragma unroll
    for (int k = 0; k < PINN_MAX_IN; ++k) {
      if (k < pi.d_in) {
        float gk[GW];
#pragma unroll
        for (int i = 0; i < GW; ++i) {
          float gg = zv0[i] * x[k];
#pragma unroll
          for (int jd = 0; jd < N1; ++jd) gg += (pi.dir1[jd] == k) ? zvd[jd][i] : 0.f;
          gk[i] = gg;
        }
        const float gs = warp_reduceg(gk, lc.lane);
        if (rlead) atomicAdd(lc.gw + g * GW + re + (long long)n1w * k, gs);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// forward of one network for the current tile, channel structure <N1, N2, PURE>.
// `phase` bit 0 = parity of the MMA barrier, bit 1 = parity of the bulk-load barrier; returned updated.
template <int N1, int N2, bool PURE, int AK>
__device__ __noinline__ uint32_t net_forward(CtaShared* cs, const DevProblem* Pp, const DevTerm* tmp, int slot,
                                             int want_grad, uint32_t phase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int C = 1 + N1 + N2;
  const DevTerm& tm = *tmp;
  const int net_id = tm.used_net[slot];
  const DevNet& net = Pp->nets[net_id];
  const DevChan& dc = tm.chan[slot];
  const TcNetSmem ns = cs->nets[net_id];
  const float* fp = reinterpret_cast<const float*>(smem + ns.fp);
  uint8_t* tP = smem + cs->off_P;
  uint8_t* tQ = smem + cs->off_Q;
  const Misc ms = misc_of(smem + cs->off_misc, cs->mx_dim, cs->mx_taps);
  const uint32_t tmem = cs->tmem;
  const bool split = cs->split != 0;
  PassInfo<N1, N2> pi;
  load_pass<N1, N2>(pi, net, dc);
  const int TL = pi.TL;
  const Tid t = tid_of();
  const int tid = t.tid, hh = t.hh, p = t.p;
  uint32_t mma_phase = phase & 1u;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < pi.d_in) ? ms.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* stash_slot = cs->stash + (size_t)slot * cs->tl_max * kTcMaxC * kTileBytes;

  dbg_mark(cs, 10);
  float u[C];                                   // last-layer partial dot products of this thread
#pragma unroll
  for (int c = 0; c < C; ++c) u[c] = 0.f;
  // the cross-warp combination of u goes through shared-memory atomics
  if (tid < C * kTcPts / 4) reinterpret_cast<float4*>(ms.scratch)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  if (kTcThreads < C * kTcPts / 4 && tid + kTcThreads < C * kTcPts / 4)
    reinterpret_cast<float4*>(ms.scratch)[tid + kTcThreads] = make_float4(0.f, 0.f, 0.f, 0.f);
  {
    // ---- layer 0 on the CUDA cores ------------------------------------------------------------------
    const int act0 = net.acts[0];
    const int ng = pi.n1w / GW;
    LoopCtx lc;
    lc.fp = tc::smem_u32(fp); lc.bt = lc.fp; lc.tP = tc::smem_u32(tP); lc.tQ = tc::smem_u32(tQ); lc.gb = nullptr; lc.gw = nullptr;
    lc.taddr = tmem + t.lane_addr; lc.act = act0; lc.split = split ? 1 : 0; lc.p = p; lc.lane = t.lane; lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH);
    lc.c0 = 0; lc.flag = (TL == 0) ? 1 : 0;
    l0_fwd_loop<N1, N2, PURE, AK>(lc, pi, x, u);
  }
  // ---- tensor layers -------------------------------------------------------------------------------------
  for (int l = 1; l <= TL; ++l) {
    const int n_in = net.dims[l], n_out = net.dims[l + 1];
    const int act = net.acts[l];
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    dbg_mark(cs, 11);
    if (tc::uni(t.warp) == 0) {
      // warp-uniform issue path: every operand is made uniform, one elected lane issues
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_Q = tc::uni(tc::smem_u32(tQ));
      const uint32_t u_whi = tc::uni(tc::smem_u32(smem + ns.w_hi[l - 1])), u_wlo = tc::uni(tc::smem_u32(smem + ns.w_lo[l - 1]));
      const int u_nk = tc::uni(n_in / 16), u_nout = tc::uni(n_out), u_split = tc::uni(split ? 1 : 0), u_wg = tc::uni(want_grad);
      const uint64_t u_stash = tc::uni((uint64_t)(stash_slot + (size_t)(l - 1) * kTcMaxC * kTileBytes));
      if (tc::elect_one()) {
        tc::tc_fence_after();
        if (u_wg) {
#pragma unroll 1
          for (int c = 0; c < C; ++c)
            tc::bulk_store_u((void*)(u_stash + (uint64_t)c * kTileBytes), u_P + c * kTileBytes, kTileBytes);
          tc::bulk_commit();
        }
        const uint32_t idesc = tc::make_idesc(128, u_nout, 0, 0);
        const uint64_t dwhi = tc::make_desc(u_whi, 0, 1024), dwlo = tc::make_desc(u_wlo, 0, 1024);
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const uint64_t dahi = tc::make_desc(u_P + c * kTileBytes, 0, 1024);
          const uint32_t d = u_tmem + TM_X + c * 64;
          mma_chain(d, dahi, dwhi, 32, 32, u_nk, idesc, 0);
          if (u_split) {
            const uint64_t dalo = tc::make_desc(u_Q + c * kTileBytes, 0, 1024);
            mma_chain(d, dahi, dwlo, 32, 32, u_nk, idesc, 1);
            mma_chain(d, dalo, dwhi, 32, 32, u_nk, idesc, 1);
          }
        }
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    dbg_mark(cs, 12);
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    dbg_mark(cs, 13);
    if (want_grad && tc::uni(t.warp) == 0) {
      if (tc::elect_one()) tc::bulk_wait_read0();       // stash copies have finished reading P (same lane issued them)
      __syncwarp();
    }
    __syncthreads();
    dbg_mark(cs, 14);
    const int ng = n_out / GW;
    LoopCtx lc;
    lc.fp = tc::smem_u32(fp); lc.bt = lc.fp + (FP_BT + (l - 1) * 64) * 4; lc.tP = tc::smem_u32(tP); lc.tQ = tc::smem_u32(tQ);
    lc.gb = nullptr; lc.gw = nullptr;
    lc.taddr = tmem + t.lane_addr; lc.act = act; lc.split = split ? 1 : 0; lc.p = p; lc.lane = t.lane;
    lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH); lc.c0 = 0; lc.flag = (l == TL) ? 1 : 0;
    tl_fwd_loop<N1, N2, PURE, AK>(lc, pi.ch, u);
  }
  // ---- last layer (n -> 1, identity): combine the column parts of every point ---------------------------------
  __syncthreads();   // scratch zeroed
  dbg_mark(cs, 15);
#pragma unroll
  for (int c = 0; c < C; ++c) atomicAdd(&ms.scratch[c * kTcPts + p], u[c]);
  __syncthreads();
  if (hh == 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) u[c] = ms.scratch[c * kTcPts + p];
    u[0] += fp[FP_BL];
    const int n_taps = tm.n_taps;
    for (int tt = 0; tt < n_taps; ++tt)
      if (tm.tap_slot[tt] == slot) {
        const int tch = tm.tap_ch[tt];
        float v = u[0];
#pragma unroll
        for (int c = 1; c < C; ++c) v = (tch == c) ? u[c] : v;
        ms.taps[tt * kTcPts + p] = v;
      }
  }
  __syncthreads();
  dbg_mark(cs, 16);
  return (phase & 2u) | mma_phase;
}

// reverse sweep of one network for the current tile
template <int N1, int N2, bool PURE, int AK>
__device__ __noinline__ uint32_t net_backward(CtaShared* cs, const DevProblem* Pp, const DevTerm* tmp, int slot,
                                              uint32_t phase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int C = 1 + N1 + N2;
  const DevTerm& tm = *tmp;
  const int net_id = tm.used_net[slot];
  const DevNet& net = Pp->nets[net_id];
  const DevChan& dc = tm.chan[slot];
  const TcNetSmem ns = cs->nets[net_id];
  const float* fp = reinterpret_cast<const float*>(smem + ns.fp);
  uint8_t* tP = smem + cs->off_P;
  uint8_t* tQ = smem + cs->off_Q;
  const Misc ms = misc_of(smem + cs->off_misc, cs->mx_dim, cs->mx_taps);
  const uint32_t tmem = cs->tmem;
  float* partial = cs->partial;
  PassInfo<N1, N2> pi;
  load_pass<N1, N2>(pi, net, dc);
  const int L = pi.L, TL = pi.TL;
  const Tid t = tid_of();
  const int tid = t.tid, hh = t.hh, p = t.p, lane = t.lane, q = t.q;
  uint32_t mma_phase = phase & 1u, ld_phase = (phase >> 1) & 1u;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < pi.d_in) ? ms.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* stash_slot = cs->stash + (size_t)slot * cs->tl_max * kTcMaxC * kTileBytes;

  dbg_mark(cs, 20);
  // adjoint of the network outputs per channel (every thread of the point needs it)
  float ub[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ub[c] = 0.f;
  {
    const int n_taps = tm.n_taps;
    for (int tt = 0; tt < n_taps; ++tt)
      if (tm.tap_slot[tt] == slot) {
        const float g = ms.tapbar[tt * kTcPts + p];
        const int tch = tm.tap_ch[tt];
#pragma unroll
        for (int c = 0; c < C; ++c) ub[c] += (tch == c) ? g : 0.f;
      }
  }
  // ---- last layer: bias gradient by warp sums; weight gradient  wbar_last[o] = sum_{c,p} ubar_c[p] H_c^{L-2}[p][o]  on the
  // tensor core: D_c[o][0..15] = H_c^T U with U[p] = (hi, lo) bf16 pairs of ubar_0..ubar_(C-1) (columns 2c, 2c+1) ------------
  {
    float* gb_last = partial + net.b_off[L - 1];
    float* gw_last = partial + net.w_off[L - 1];
    if (hh == 0) {
      const float s = warp_sum<float>(ub[0]);
      if (lane == 0) atomicAdd(gb_last, s);
      uint32_t w[8];
#pragma unroll
      for (int c = 0; c < 8; ++c) w[c] = 0u;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t hi = tc::pack_bf16(ub[c], 0.f) & 0xffffu;
        const float r = ub[c] - __uint_as_float(hi << 16);
        w[c] = hi | (tc::pack_bf16(r, 0.f) << 16);
      }
      const uint32_t q0 = tc::smem_u32(tQ);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + tc::swz_chunk(p, 0)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + tc::swz_chunk(p, 1)), "r"(w[4]), "r"(w[5]), "r"(w[6]), "r"(w[7]) : "memory");
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_Q = tc::uni(tc::smem_u32(tQ));
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc(128, 16, 1, 1);
        const uint64_t db = tc::make_desc(u_Q, 0, 1024);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          mma_chain(u_tmem + TM_Y + 16 * c, tc::make_desc(u_P + c * kTileBytes, 0, 1024), db, 2048, 2048, kTcPts / 16, idesc, 0);
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    if (q < 2 && hh == 0) {
      const int o = q * 32 + lane;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v[2];
        tmem_ld2(tmem + t.lane_addr + TM_Y + 16 * c + 2 * c, v);
        tc::tmem_ld_wait();
        acc += v[0] + v[1];
      }
      if (o < pi.nL) atomicAdd(gw_last + o, acc);
    }
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
  }

  // ---- tensor layers, last to first ------------------------------------------------------------------------------------
  for (int l = TL; l >= 1; --l) {
    const int n_in = net.dims[l], n_out = net.dims[l + 1];
    const int act = net.acts[l];
    float* gb = partial + net.b_off[l];
    float* gw = partial + net.w_off[l];
    const float* bt = fp + FP_BT + (l - 1) * 64;
    const uint32_t whi = tc::smem_u32(smem + ns.w_hi[l - 1]);
    dbg_mark(cs, 21);
    if (tc::uni(t.warp) == 0) {
      // reload this layer's input tiles H^{l-1} (bf16 hi) from the stash into Q
      const uint32_t u_Q = tc::uni(tc::smem_u32(tQ)), u_bar = tc::uni(tc::smem_u32(ms.bar_ld));
      const uint64_t u_stash = tc::uni((uint64_t)(stash_slot + (size_t)(l - 1) * kTcMaxC * kTileBytes));
      if (tc::elect_one()) {
        tc::mbar_arrive_expect_tx_u(u_bar, C * kTileBytes);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          tc::bulk_load_u(u_Q + c * kTileBytes, (const void*)(u_stash + (uint64_t)c * kTileBytes), kTileBytes, u_bar);
      }
      __syncwarp();
    }
    wait_bar(ms.bar_ld, ld_phase);
    dbg_mark(cs, 22);
    // recompute pre-activations in groups of <= 32 columns and turn output adjoints into Zbar tiles
    for (int c0 = 0; c0 < n_out; c0 += 32) {
      const int gw_cols = (n_out - c0) < 32 ? (n_out - c0) : 32;     // 32 or 16
      tc::tc_fence_before();
      __syncthreads();
      dbg_mark(cs, 23);
      if (tc::uni(t.warp) == 0) {
        const uint32_t u_tmem = tc::uni(tmem), u_Q = tc::uni(tc::smem_u32(tQ)), u_w = tc::uni(whi + c0 * 128);
        const int u_nk = tc::uni(n_in / 16), u_gw = tc::uni(gw_cols);
        if (tc::elect_one()) {
          tc::tc_fence_after();
          const uint32_t idesc = tc::make_idesc(128, u_gw, 0, 0);
          const uint64_t dw = tc::make_desc(u_w, 0, 1024);
#pragma unroll 1
          for (int c = 0; c < C; ++c)
            mma_chain(u_tmem + TM_Y + c * 32, tc::make_desc(u_Q + c * kTileBytes, 0, 1024), dw, 32, 32, u_nk, idesc, 0);
          tc::mma_commit(ms.bar_mma);
        }
        __syncwarp();
      }
      dbg_mark(cs, 24);
      wait_bar(ms.bar_mma, mma_phase);
      tc::tc_fence_after();
      dbg_mark(cs, 25);
      const int ng = gw_cols / GWB;
      LoopCtx lc;
      lc.fp = tc::smem_u32(fp); lc.bt = tc::smem_u32(bt); lc.tP = tc::smem_u32(tP); lc.tQ = tc::smem_u32(tQ); lc.gb = gb; lc.gw = nullptr;
      lc.taddr = tmem + t.lane_addr;
      lc.act = act; lc.split = 0; lc.p = p; lc.lane = lane; lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH);
      lc.c0 = c0; lc.flag = (l == TL) ? 1 : 0;
      tl_bwd_loop<N1, N2, PURE, AK>(lc, pi.ch, ub);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    dbg_mark(cs, 26);
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_Q = tc::uni(tc::smem_u32(tQ)), u_w = tc::uni(whi);
      const uint32_t u_ones = tc::uni(tc::smem_u32(smem + cs->off_ones));
      const int u_nin = tc::uni(n_in), u_nko = tc::uni(n_out / 16);
      if (tc::elect_one()) {
        tc::tc_fence_after();
        // dgrad: Hbar_c = Zbar_c * W_l  -> X
        const uint32_t idg = tc::make_idesc(128, u_nin, 0, 1);
        const uint64_t dw = tc::make_desc(u_w, 0, 1024);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          mma_chain(u_tmem + TM_X + c * 64, tc::make_desc(u_P + c * kTileBytes, 0, 1024), dw, 32, 2048, u_nko, idg, 0);
        // wgrad: Wbar_l = sum_c Zbar_c^T * H_c -> Y (rows >= 64 alias rows - 64 through LBO = 0)
        const uint32_t iwg = tc::make_idesc(128, u_nin, 1, 1);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          mma_chain(u_tmem + TM_Y, tc::make_desc(u_P + c * kTileBytes, 0, 1024), tc::make_desc(u_Q + c * kTileBytes, 0, 1024),
                    2048, 2048, kTcPts / 16, iwg, c > 0 ? 1u : 0u);
        // bias gradient: bbar_l[o] = sum_p Zbar_0[p][o] -> Y column 64 (B = the constant ones atom, SBO = 0, no k advance)
        mma_chain(u_tmem + TM_Y + 64, tc::make_desc(u_P, 0, 1024), tc::make_desc(u_ones, 0, 0), 2048, 0, kTcPts / 16,
                  tc::make_idesc(128, 16, 1, 1), 0);
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    dbg_mark(cs, 27);
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    dbg_mark(cs, 28);
    // flush the weight-gradient tile: TMEM lane = output neuron o, column = input neuron k
    if (q < 2) {
      const int o = q * 32 + lane;
      if (hh == 0) {
        float v[2];
        tmem_ld2(tmem + t.lane_addr + TM_Y + 64, v);
        tc::tmem_ld_wait();
        if (o < n_out) atomicAdd(gb + o, v[0]);
      }
      const int part = n_in / kNH;
#pragma unroll 1
      for (int k0 = hh * part; k0 < (hh + 1) * part; k0 += 4) {
        float v[4];
        tmem_ld4(tmem + t.lane_addr + TM_Y + k0, v);
        tc::tmem_ld_wait();
        if (o < n_out) {
#pragma unroll
          for (int i = 0; i < 4; ++i) atomicAdd(gw + o + (long long)n_out * (k0 + i), v[i]);
        }
      }
    }
  }

  // ---- layer 0 backward ---------------------------------------------------------------------------------------------------
  {
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    dbg_mark(cs, 29);
    const int act0 = net.acts[0];
    float* gb0 = partial + net.b_off[0];
    float* gw0 = partial + net.w_off[0];
    LoopCtx lc;
    lc.fp = tc::smem_u32(fp); lc.bt = lc.fp; lc.tP = tc::smem_u32(tP); lc.tQ = tc::smem_u32(tQ); lc.gb = gb0; lc.gw = gw0;
    lc.taddr = tmem + t.lane_addr; lc.act = act0; lc.split = 0; lc.p = p; lc.lane = lane; lc.c0 = 0; lc.flag = (TL == 0) ? 1 : 0;
    if (TL == 0) {
      // no tensor layer: everything on the CUDA cores (warp reduce-scatter + atomics)
      const int ng = pi.n1w / GW;
      lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH);
      l0_bwd_loop<N1, N2, PURE, AK>(lc, pi, x, ub);
    } else {
      // Wbar_0[o][k] = sum_p zbar_0[p][o] x_k[p] + zbar_(1+j)[p][o] [dir1[j] == k],  bbar_0[o] = sum_p zbar_0[p][o]:
      // one MMA chain  D[o][0..15] = Zbar_0^T [x | 1] + sum_j Zbar_(1+j)^T E_(dir1[j])  with K = the 128 points.
      // B tiles (rows = points, 16 columns used) live in Q, which is free after the last tensor layer:
      //   tile 0: bf16 hi of (x_0..x_7) in columns 0..7, 1.0 in column 8;  tile 1+j: 1.0 in column dir1[j];
      //   tile 1+N1: bf16 lo of x (only when the channel count leaves a spare tile, i.e. N2 > 0)
      if (tid < kTcPts) {
        uint32_t hi[4], lo[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          hi[k] = tc::pack_bf16(x[2 * k], x[2 * k + 1]);
          lo[k] = tc::pack_bf16(x[2 * k] - __uint_as_float(hi[k] << 16), x[2 * k + 1] - __uint_as_float(hi[k] & 0xffff0000u));
        }
        const uint32_t q0 = tc::smem_u32(tQ);
        const uint32_t c0a = tc::swz_chunk(p, 0), c1a = tc::swz_chunk(p, 1);
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + c0a), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + c1a), "r"(0x00003f80u), "r"(0u), "r"(0u), "r"(0u) : "memory");
        if (N2 > 0) {
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + (1 + N1) * kTileBytes + c0a), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + (1 + N1) * kTileBytes + c1a), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
        }
#pragma unroll
        for (int j = 0; j < N1; ++j) {
          const int d = pi.dir1[j];
          uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
          for (int k = 0; k < 4; ++k) w[k] = (d == 2 * k) ? 0x00003f80u : ((d == 2 * k + 1) ? 0x3f800000u : 0u);
          const uint32_t qb = q0 + (1 + j) * kTileBytes;
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(qb + c0a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(qb + c1a), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
        }
      }
      const int ng = pi.n1w / GWB;
      lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH);
      l0_bwd_store_loop<N1, N2, PURE, AK>(lc, pi, x);
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (tc::uni(t.warp) == 0) {
        const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_Q = tc::uni(tc::smem_u32(tQ));
        if (tc::elect_one()) {
          tc::tc_fence_after();
          const uint32_t idesc = tc::make_idesc(128, 16, 1, 1);
          const uint64_t a0 = tc::make_desc(u_P, 0, 1024);
          mma_chain(u_tmem + TM_Y, a0, tc::make_desc(u_Q, 0, 1024), 2048, 2048, kTcPts / 16, idesc, 0);
          if (N2 > 0)
            mma_chain(u_tmem + TM_Y, a0, tc::make_desc(u_Q + (1 + N1) * kTileBytes, 0, 1024), 2048, 2048, kTcPts / 16, idesc, 1);
#pragma unroll 1
          for (int j = 0; j < N1; ++j)
            mma_chain(u_tmem + TM_Y, tc::make_desc(u_P + (1 + j) * kTileBytes, 0, 1024),
                      tc::make_desc(u_Q + (1 + j) * kTileBytes, 0, 1024), 2048, 2048, kTcPts / 16, idesc, 1);
          tc::mma_commit(ms.bar_mma);
        }
        __syncwarp();
      }
      wait_bar(ms.bar_mma, mma_phase);
      tc::tc_fence_after();
      if (q < 2 && hh == 0) {
        const int o = q * 32 + lane;
        float v[16];
        tc::tmem_ld16(tmem + t.lane_addr + TM_Y, v);
        tc::tmem_ld_wait();
        if (o < pi.n1w) {
#pragma unroll
          for (int k = 0; k < PINN_MAX_IN; ++k)
            if (k < pi.d_in) atomicAdd(gw0 + o + (long long)pi.n1w * k, v[k]);
          atomicAdd(gb0 + o, v[8]);
        }
      }
    }
  }
  __syncthreads();
  dbg_mark(cs, 30);
  return (ld_phase << 1) | mma_phase;
}


__global__ void __launch_bounds__(kTcThreads, 1) tc_loss_grad_kernel(const __grid_constant__ TcArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ CtaShared cs;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const DevProblem* Pp = args.prob;
  const DevProblem& P = *Pp;
  const Misc ms = misc_of(smem + args.off_misc, args.mx_dim, args.mx_taps);
  float* partial = args.partial + (long long)blockIdx.x * args.partial_stride;
  const bool want_grad = (args.mode == 0);
  const float* theta = args.theta;
#ifdef PINN_DEBUG
  long long span_c0 = 0;
  unsigned long long span_g0 = 0;
  if (args.dbg && tid == 0) {
    span_c0 = clock64();
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(span_g0));
  }
#endif

  // ---- per-CTA setup --------------------------------------------------------------------------------------------------------
  // The step usually starts with everything cold in L2 (the caller's other work evicted it).  Pull what the serial setup
  // code is about to read -- theta (tens of KB), the network / term descriptors and this CTA's first point tile -- into L2
  // right away, from counts that travel in the launch arguments: the dependent loads below then pay L2 latency, not a chain
  // of HBM round trips.
  for (long long i = (long long)tid * 32; i < args.n_theta; i += (long long)kTcThreads * 32) tc::prefetch_l2(theta + i);
  {
    const char* nets0 = reinterpret_cast<const char*>(&Pp->nets[0]);
    const char* terms0 = reinterpret_cast<const char*>(&Pp->terms[0]);
    const int nb_nets = args.n_nets * (int)sizeof(DevNet), nb_terms = args.n_terms * (int)sizeof(DevTerm);
    if (tid == 0) tc::prefetch_l2(Pp);
    for (int o = tid * 128; o < nb_nets; o += kTcThreads * 128) tc::prefetch_l2(nets0 + o);
    for (int o = tid * 128; o < nb_terms; o += kTcThreads * 128) tc::prefetch_l2(terms0 + o);
    const int tile = args.tile_begin + (int)blockIdx.x;
    if (tile < args.tile_end && tid < 32) {
      int ti = 0;
      while (ti + 1 < args.n_terms && tile >= args.dyn[ti + 1].tile0) ++ti;
      const long long p0 = (long long)(tile - args.dyn[ti].tile0) * kTcPts;
      const long long row_bytes = 4ll * args.term_dim[ti];
      const long long off = p0 * row_bytes + (long long)tid * 128;
      if (off < args.dyn[ti].n * row_bytes) tc::prefetch_l2(reinterpret_cast<const char*>(args.dyn[ti].pts) + off);
    }
  }
  if (tid == 0) {
    tc::mbar_init(ms.bar_mma, 1);
    tc::mbar_init(ms.bar_ld, 1);
    tc::fence_barrier_init();
    cs.split = args.split; cs.tl_max = args.tl_max; cs.off_P = args.off_P; cs.off_Q = args.off_Q;
    cs.off_misc = args.off_misc; cs.off_ones = args.off_ones; cs.mx_dim = args.mx_dim; cs.mx_taps = args.mx_taps;
    cs.partial = partial;
    cs.stash = args.stash + (long long)blockIdx.x * args.stash_per_cta;
    cs.theta = theta;
#ifdef PINN_DEBUG
    cs.dbg = (blockIdx.x == 0) ? args.dbg : nullptr;
#else
    cs.dbg = nullptr;
#endif
    cs.dbg_n = 0;
    for (int k = 0; k < PINN_MAX_NETS; ++k) cs.nets[k] = args.nets[k];
#ifdef PINN_DEBUG
    if (cs.dbg) cs.dbg[cs.dbg_n++] = ((long long)1 << 48) | (clock64() & 0xffffffffffffLL);
#endif
  }
  if (warp == 0) tc::tmem_alloc<512>(ms.tmem_slot);
  if (want_grad) {
    const long long n4 = P.n_theta / 4;
    float4* p4 = reinterpret_cast<float4*>(partial);
    if ((reinterpret_cast<uintptr_t>(partial) & 15) == 0) {
      for (long long i = tid; i < n4; i += kTcThreads) p4[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      for (long long i = n4 * 4 + tid; i < P.n_theta; i += kTcThreads) partial[i] = 0.f;
    } else {
      for (long long i = tid; i < P.n_theta; i += kTcThreads) partial[i] = 0.f;
    }
  }
  if (tid < PINN_MAX_TERMS) ms.tsum[tid] = 0.0;
  if (tid < 64) {      // ones atom: row r (128 B) holds bf16 1.0 in logical column 0 = 16-byte chunk (0 ^ r)
    const int r = tid >> 3, ch = tid & 7;
    *reinterpret_cast<uint4*>(smem + args.off_ones + r * 128 + ch * 16) = make_uint4(ch == r ? 0x00003f80u : 0u, 0u, 0u, 0u);
  }
  // stage weights: bf16 hi / lo operand tiles of the tensor layers, fp32 blocks of the first / last layers
  for (int kn = 0; kn < P.n_nets; ++kn) {
    const DevNet& net = P.nets[kn];
    const TcNetSmem& ns = args.nets[kn];
    if (ns.fp < 0) continue;
    float* fp = reinterpret_cast<float*>(smem + ns.fp);
    const int L = net.n_layers;
    for (int i = tid; i < FP_SIZE; i += kTcThreads) fp[i] = 0.f;
    __syncthreads();
    const int n1w = net.dims[1], d_in = net.dims[0];
    const long long w0 = net.w_off[0], b0 = net.b_off[0];
    for (int i = tid; i < n1w * d_in; i += kTcThreads) {
      const int o = i % n1w, k = i / n1w;
      fp[FP_W1 + o * 8 + k] = __ldg(&theta[w0 + i]);
    }
    for (int i = tid; i < n1w; i += kTcThreads) fp[FP_B1 + i] = __ldg(&theta[b0 + i]);
    // thread <-> (layer, row o, chunk of 8 k).  All global loads of all layers are issued before the first conversion
    // (fully unrolled, predicated on the layer count): one memory round trip instead of one per layer -- the step
    // starts with theta cold in L2 when the caller's other work has evicted it
    {
      constexpr int kItMax = kTcMaxTL * 512 / kTcThreads;
      const int n_items = (L - 2) * 512;
      float w[kItMax][8];
#pragma unroll
      for (int it = 0; it < kItMax; ++it) {
        const int i = tid + it * kTcThreads;
        if (i < n_items) {
          const int l = 1 + i / 512, r = i & 511;
          const int o = r & 63, kc = r >> 6;
          const int n_in = net.dims[l], n_out = net.dims[l + 1];
          const long long woff = net.w_off[l];
#pragma unroll
          for (int e = 0; e < 8; ++e) {
            const int k = kc * 8 + e;
            w[it][e] = (o < n_out && k < n_in) ? __ldg(&theta[woff + o + (long long)n_out * k]) : 0.f;
          }
        }
      }
#pragma unroll
      for (int it = 0; it < kItMax; ++it) {
        const int i = tid + it * kTcThreads;
        if (i < n_items) {
          const int l = 1 + i / 512, r = i & 511;
          const int o = r & 63, kc = r >> 6;
          uint8_t* thi = smem + ns.w_hi[l - 1];
          uint8_t* tlo = smem + ns.w_lo[l - 1];
          uint4 h, lo4;
          h.x = tc::pack_bf16(w[it][0], w[it][1]); h.y = tc::pack_bf16(w[it][2], w[it][3]);
          h.z = tc::pack_bf16(w[it][4], w[it][5]); h.w = tc::pack_bf16(w[it][6], w[it][7]);
          *reinterpret_cast<uint4*>(thi + tc::swz_chunk(o, kc)) = h;
          if (args.split) {
            lo4.x = tc::pack_bf16(w[it][0] - __uint_as_float(h.x << 16), w[it][1] - __uint_as_float(h.x & 0xffff0000u));
            lo4.y = tc::pack_bf16(w[it][2] - __uint_as_float(h.y << 16), w[it][3] - __uint_as_float(h.y & 0xffff0000u));
            lo4.z = tc::pack_bf16(w[it][4] - __uint_as_float(h.z << 16), w[it][5] - __uint_as_float(h.z & 0xffff0000u));
            lo4.w = tc::pack_bf16(w[it][6] - __uint_as_float(h.w << 16), w[it][7] - __uint_as_float(h.w & 0xffff0000u));
            *reinterpret_cast<uint4*>(tlo + tc::swz_chunk(o, kc)) = lo4;
          }
        }
      }
    }
    for (int i = tid; i < (L - 2) * 64; i += kTcThreads) {
      const int l = 1 + i / 64, o = i & 63;
      if (o < net.dims[l + 1]) fp[FP_BT + (l - 1) * 64 + o] = __ldg(&theta[net.b_off[l] + o]);
    }
    const int nL = net.dims[L - 1];
    const long long wl = net.w_off[L - 1], bl = net.b_off[L - 1];
    for (int i = tid; i < nL; i += kTcThreads) fp[FP_WL + i] = __ldg(&theta[wl + i]);
    if (tid == 0) fp[FP_BL] = __ldg(&theta[bl]);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) cs.tmem = *ms.tmem_slot;
  __syncthreads();
  dbg_mark(&cs, 2);
  uint32_t phase = 0;

  for (int tile = args.tile_begin + blockIdx.x; tile < args.tile_end; tile += gridDim.x) {
    int ti = 0;
    while (ti + 1 < P.n_terms && tile >= args.dyn[ti + 1].tile0) ++ti;
    const DevTerm* tmp = &P.terms[ti];
    const DevTerm& tm = *tmp;
    const long long p0 = (long long)(tile - args.dyn[ti].tile0) * kTcPts;
    const long long n_pts = args.dyn[ti].n;
    const float* pts = reinterpret_cast<const float*>(args.dyn[ti].pts);
    const float* qw = reinterpret_cast<const float*>(args.dyn[ti].qw);
    // warm L1 with the term header + residual program and the network descriptors (read by every phase)
    if (tid < (int)((sizeof(DevTerm) + 127) / 128)) tc::prefetch_l1(reinterpret_cast<const char*>(tmp) + tid * 128);
    if (tid >= 128 && tid < 128 + (int)((sizeof(DevNet) * PINN_MAX_NETS + 127) / 128))
      tc::prefetch_l1(reinterpret_cast<const char*>(&P.nets[0]) + (tid - 128) * 128);
    const int dim = tm.dim, n_taps = tm.n_taps, n_used = tm.n_used, weighted = tm.weighted;
    // Collocation tile: one point = dim contiguous scalars (the reference's d x N train-set layout), so a full 128-point
    // tile is ONE contiguous block of dim x 512 bytes.  The TMA unit copies it into shared memory in a single bulk transfer
    // (cp.async.bulk + mbarrier transaction count; staged in the scratch array, free between tiles) and 128 threads
    // transpose it to [row][point].  Partial last tiles (clamped rows) and callers' buffers that are not 16-byte aligned
    // take the per-element path.
    const float* tile_src = pts + p0 * dim;
    const bool bulk_tile = (p0 + kTcPts <= n_pts) && dim <= kTcMaxC && ((reinterpret_cast<uintptr_t>(tile_src) & 15) == 0);
    if (bulk_tile) {
      uint32_t ldp = (phase >> 1) & 1u;
      if (tid == 0) {
        tc::mbar_arrive_expect_tx(ms.bar_ld, (uint32_t)(dim * kTcPts * 4));
        tc::bulk_load(ms.scratch, tile_src, (uint32_t)(dim * kTcPts * 4), ms.bar_ld);
      }
      wait_bar(ms.bar_ld, ldp);
      phase = (phase & 1u) | (ldp << 1);
      if (tid < kTcPts)
        for (int r = 0; r < dim; ++r) ms.Xs[r * kTcPts + tid] = ms.scratch[tid * dim + r];
    } else {
      for (int i = tid; i < dim * kTcPts; i += kTcThreads) {
        int pp = i / dim, r = i - pp * dim;
        long long gp = p0 + pp;
        if (gp >= n_pts) gp = n_pts - 1;
        ms.Xs[r * kTcPts + pp] = pts[gp * dim + r];
      }
    }
    if (tid < kTcPts) {
      long long gp = p0 + tid;
      float w = 0.f;
      if (gp < n_pts) w = weighted ? qw[gp] : 1.f;
      ms.qws[tid] = w;
    }
    for (int i = tid; i < n_taps * kTcPts; i += kTcThreads) ms.tapbar[i] = 0.f;
    __syncthreads();
    dbg_mark(&cs, 3);

    for (int slot = 0; slot < n_used; ++slot) {
      const int k1 = tm.chan[slot].n1, k2 = tm.chan[slot].n2, pu = tm.chan[slot].pure;
      const int ak = args.net_ak[tm.used_net[slot]];
      PINN_TC_DISPATCH(k1, k2, pu, ak, (phase = net_forward<A1, A2, PU, AK>(&cs, Pp, tmp, slot, want_grad ? 1 : 0, phase)));
    }

    dbg_mark(&cs, 4);
    // the lo-tile region Q is free between the forward and the reverse sweep: stage the program text and the
    // per-point value / adjoint arrays there (shared memory instead of global + local memory)
    const int n_instr = tm.n_instr;
    DevInstr* sprog = reinterpret_cast<DevInstr*>(smem + args.off_Q);
    float* sval = reinterpret_cast<float*>(smem + args.off_Q + 8192);
    const bool prog_sm = (size_t)8192 + (size_t)2 * n_instr * kTcPts * 4 <= (size_t)(args.off_Q_bytes);
    if (prog_sm) {
      const int nw = n_instr * (int)(sizeof(DevInstr) / 4);
      const int* src = reinterpret_cast<const int*>(tm.prog);
      for (int i = tid; i < nw; i += kTcThreads) reinterpret_cast<int*>(sprog)[i] = __ldg(src + i);
      __syncthreads();
    }
    // ---- residual program, loss partial, tap adjoints (threads 0..127: one point each) -----------------------------------------
    if (tid < kTcPts) {
      float pbar[PINN_MAX_PARAMS];
#pragma unroll
      for (int j = 0; j < PINN_MAX_PARAMS; ++j) pbar[j] = 0.f;
      float r;
      if (prog_sm) {
        r = run_program_t<float, kTcPts, true>(sprog, n_instr, theta + P.param_off, ms.Xs, ms.taps, ms.tapbar, pbar, tid,
                                               want_grad, sval, sval + n_instr * kTcPts);
      } else {
        r = run_program<float, kTcPts>(tm, theta + P.param_off, ms.Xs, ms.taps, ms.tapbar, pbar, tid, want_grad);
      }
      const float w = ms.qws[tid];
      double s = (double)w * (double)r * (double)r;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) atomicAdd(&ms.tsum[ti], s);
      if (args.mode == 2) {
        long long gp = p0 + tid;
        if (gp < n_pts) args.resid_out[gp] = r;
      }
      if (want_grad) {
        const float g = (float)args.seed[ti] * w * 2.f * r;
        for (int tt = 0; tt < n_taps; ++tt) ms.tapbar[tt * kTcPts + tid] *= g;
        const int n_params = P.n_params;
        for (int j = 0; j < n_params; ++j) {
          float v = warp_sum<float>(pbar[j] * g);
          if (lane == 0) atomicAdd(&partial[P.param_off + j], v);
        }
      }
    }
    __syncthreads();

    dbg_mark(&cs, 5);
    if (want_grad) {
      if (tid == 0) tc::bulk_wait0();             // stash writes of this tile are complete before reloads
      __syncthreads();
      dbg_mark(&cs, 6);
      for (int slot = n_used - 1; slot >= 0; --slot) {
        const int k1 = tm.chan[slot].n1, k2 = tm.chan[slot].n2, pu = tm.chan[slot].pure;
        const int ak = args.net_ak[tm.used_net[slot]];
        if (n_used > 1) {
          // P must hold this slot's last hidden activations again: recompute its forward
          PINN_TC_DISPATCH(k1, k2, pu, ak, (phase = net_forward<A1, A2, PU, AK>(&cs, Pp, tmp, slot, 0, phase)));
        }
        PINN_TC_DISPATCH(k1, k2, pu, ak, (phase = net_backward<A1, A2, PU, AK>(&cs, Pp, tmp, slot, phase)));
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  dbg_mark(&cs, 7);
#ifdef PINN_DEBUG
  if (tid == 0 && cs.dbg) cs.dbg[999] = cs.dbg_n;
  if (args.dbg && tid == 0 && blockIdx.x < 250) {
    unsigned long long g1;
    unsigned int smid;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
    asm volatile("mov.u32 %0, %smid;" : "=r"(smid));
    long long* rec = args.dbg + 1000 + 4 * blockIdx.x;
    rec[0] = (long long)span_g0; rec[1] = (long long)g1; rec[2] = clock64() - span_c0; rec[3] = smid;
  }
#endif
  if (tid < PINN_MAX_TERMS) args.term_sums[(long long)blockIdx.x * PINN_MAX_TERMS + tid] = ms.tsum[tid];
  if (warp == 0) tc::tmem_dealloc<512>(cs.tmem);
  // gradient reduction, optimizer step and the multi-GPU sum in the kernel tail (tail.cuh)
  if (args.tail.state)
    fused_tail<float, kTcThreads>(args.tail, args.partial, args.partial_stride, args.term_sums, P.n_theta, P.n_terms, want_grad ? 1 : 0,
                                  reinterpret_cast<float*>(smem + args.off_P));
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
size_t tc_misc_bytes(int mx_dim, int mx_taps) {
  return (size_t)mx_dim * kTcPts * 4 + 2 * (size_t)mx_taps * kTcPts * 4 + (size_t)kTcMaxC * kTcPts * 4 + kTcPts * 4 +
         PINN_MAX_TERMS * 8 + 16 + 16;
}

cudaError_t tc_launch(const TcArgs& a, int grid, size_t smem, cudaStream_t st) {
  static size_t granted[64] = {0};
  cudaError_t e = ensure_dynamic_smem(tc_loss_grad_kernel, smem, granted);
  if (e != cudaSuccess) return e;
  return launch_fused_kernel(tc_loss_grad_kernel, a, grid, kTcThreads, smem, st, a.tail.state != nullptr);
}

}  // namespace pinn
