// tc_wide_kernel.cu -- fused PINN loss+gradient kernel, tcgen05 path for 128-wide layers (sm_100a, bf16 operands).
//
// Same decomposition as tc_kernel.cu (one CTA of 512 threads per 128-point tile, a point is a TMEM lane and a row
// of every operand tile, derivative channels share the weight operand), re-planned for hidden widths 64 / 128 where
// neither the weights of all layers nor two generations of activations fit in shared memory:
//
//   * an activation set is C channels x 2 tiles (128 points x 64 bf16, 128-byte swizzle) = C x 32 KB in region P;
//     TMEM holds C x 128 accumulator columns (C <= 4), so the epilogue of a layer overwrites its own input in place;
//   * weights are packed once per step by tw_pack_kernel into bf16 swizzled images (32 KB per layer) and streamed
//     through two 32 KB buffers S0 / S1 with cp.async.bulk + mbarrier, prefetched one layer ahead;
//   * the forward sweep stashes every tensor layer's input tiles (bf16, for wgrad) and biased pre-activations
//     (fp32, point-fastest so that a warp writes / reads 256 contiguous bytes) to a per-CTA global buffer (L2);
//     the reverse sweep reads the pre-activations straight into registers -- no recompute, no TMEM for it;
//   * reverse, per tensor layer:  Zbar tiles -> P;  wgrad  Wbar_l = sum_c Zbar_c^T H_c  with H_c streamed through
//     S0 / S1 per channel;  then dgrad  Hbar_c = Zbar_c W_l  with W_l in the buffer wgrad released first.
//     MN-major operands whose M / N extent is 128 span two tiles through the descriptor's leading-dimension byte
//     offset (pinned on hardware by scripts/tc_probe_wide.py).
//
// Replaces the same reference functions as the other paths (Phi src/pinn_types.jl:79-90, numeric_derivative
// :445-482, the generated residual and mean(abs2) src/training_strategies.jl:215-221, Zygote gradient
// src/discretize.jl:778).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "tc_common.cuh"
#include "tail.cuh"

namespace pinn {

constexpr uint32_t TB = kTileBytes;

struct TwShared {
  uint32_t tmem;
  int tl_max, off_P, off_S, off_misc, off_ones, off_nets, mx_dim, mx_taps;
  float* partial;
  uint8_t* hstash;
  float* zstash;
  const float* theta;
  const uint8_t* wpack;
  long long* dbg;
  int dbg_n;
  int off_fp[PINN_MAX_NETS], wimg[PINN_MAX_NETS];
  int next_tile;                     // dynamic scheduler: tile claimed for the next iteration
  uint32_t ph_ld[2], ph_free[2];     // phases of the streaming barriers, owned by the issuing lane of warp 0
  uint64_t bar_ld[2], bar_free[2];   // S0 / S1: bytes landed, MMAs that read the buffer retired
};

// first-layer pre-activations of neuron o (channel vector zz); fpa = shared-memory address of the fp32 block
template <int N1, int N2>
__device__ __forceinline__ void first_layer_elem_w(uint32_t fpa, const PassInfo<N1, N2>& pi, const float (&x)[PINN_MAX_IN],
                                                   int o, float* zz) {
  float s = lds_f32(fpa + (FW_B1 + o) * 4);
  const uint32_t wa = fpa + (FW_W1 + o * 8) * 4;
  if (pi.d_in <= 3) {
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
  for (int j = 0; j < N1; ++j) zz[1 + j] = lds_f32(fpa + (FW_W1 + o * 8 + pi.dir1[j]) * 4);
#pragma unroll
  for (int j = 0; j < N2; ++j) zz[1 + N1 + j] = 0.f;
}

struct LoopW {
  uint32_t fp;            // shared-memory address of the network's fp32 parameter block
  uint32_t bt;            // shared-memory address of the current tensor layer's bias
  uint32_t tP;            // shared-memory address of the operand tiles (channel c, column block kb: (c*2+kb)*TB)
  float* gb;              // bias gradient of the current layer (CTA partial)
  uint32_t taddr;         // tmem base + lane quadrant
  int act, p, lane, g0, g1, flag;
};

__device__ __forceinline__ uint32_t tile_of(uint32_t tP, int c, int col) { return tP + (uint32_t)(c * 2 + (col >> 6)) * TB; }

// layer 0 forward: coordinates -> H^0 tiles
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tw_l0_fwd_loop(const LoopW lc, const PassInfo<N1, N2> pi, const float* xp) {
  constexpr int C = 1 + N1 + N2;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = xp[k];
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float h[C][4];
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      float za[C], zb2[C];
      first_layer_elem_w<N1, N2>(lc.fp, pi, x, g * 4 + i, za);
      first_layer_elem_w<N1, N2>(lc.fp, pi, x, g * 4 + i + 1, zb2);
      P2 zz[C], hv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) zz[c] = mk2(za[c], zb2[c]);
      chain_fwd<N1, N2, PURE, AK, P2>(lc.act, pi.ch, zz, hv);
#pragma unroll
      for (int c = 0; c < C; ++c) { h[c][i] = hv[c].v.x; h[c][i + 1] = hv[c].v.y; }
    }
    const int col = g * 4;
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(tile_of(lc.tP, c, col), 0u, lc.p, col & 63, h[c], false);
  }
}

// tensor layer forward epilogue: TMEM accumulators -> bias + activation chain -> next operand tiles (in place),
// biased pre-activations -> fp32 stash (zst != nullptr), last-layer dot products (flag)
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tw_fwd_loop(const LoopW lc, const Chan<N1, N2> ch, float* up, float2* zst) {
  constexpr int C = 1 + N1 + N2;
  float u[C];
#pragma unroll
  for (int c = 0; c < C; ++c) u[c] = up[c];
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    float z[C][4];
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ld4(lc.taddr + c * kTwW + g * 4, z[c]);
    tc::tmem_ld_wait();
    const int col = g * 4;
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      P2 zz[C], hv[C];
      zz[0] = mk2(z[0][i] + lds_f32(lc.bt + (col + i) * 4), z[0][i + 1] + lds_f32(lc.bt + (col + i + 1) * 4));
#pragma unroll
      for (int c = 1; c < C; ++c) zz[c] = mk2(z[c][i], z[c][i + 1]);
      if (zst) {
#pragma unroll
        for (int c = 0; c < C; ++c) zst[(c * 64 + ((col + i) >> 1)) * kTcPts] = zz[c].v;
      }
      chain_fwd<N1, N2, PURE, AK, P2>(lc.act, ch, zz, hv);
#pragma unroll
      for (int c = 0; c < C; ++c) { z[c][i] = hv[c].v.x; z[c][i + 1] = hv[c].v.y; }
      if (lc.flag) {
        const float w0 = lds_f32(lc.fp + (FW_WL + col + i) * 4), w1 = lds_f32(lc.fp + (FW_WL + col + i + 1) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) u[c] = fmaf(w1, hv[c].v.y, fmaf(w0, hv[c].v.x, u[c]));
      }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(tile_of(lc.tP, c, col), 0u, lc.p, col & 63, z[c], false);
  }
#pragma unroll
  for (int c = 0; c < C; ++c) up[c] = u[c];
}

// tensor layer reverse epilogue: stashed pre-activations and output adjoints (TMEM X, or w_last * ubar for the last
// hidden layer: flag) -> Zbar tiles in P (the bias gradient is a column sum of Zbar_0, taken by an MMA chain against
// the ones atom).  The stash is read from HBM more often than from L2 (148 CTAs x 1.5 MB in flight > L2); the next
// granule's loads are in flight while the current one is processed.  Measured on the same box (profiles/
// r01_wide_timeline.md): no prefetch 0.652 ms, this 0.648 ms, two granules ahead 0.737 ms (spills), bf16 stash of the
// derivative channels 0.669 ms and 4x the gradient error -- the phase is bound by the burst of HBM reads, not by latency.
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tw_bwd_loop(const LoopW lc, const Chan<N1, N2> ch, const float* ubp, const float2* zst) {
  constexpr int C = 1 + N1 + N2;
  float ub[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ub[c] = ubp[c];
  float2 za[C][2], zb[C][2];
  auto load = [&](float2 (&zz)[C][2], int g) {
#pragma unroll
    for (int c = 0; c < C; ++c) {
      zz[c][0] = zst[(c * 64 + 2 * g) * kTcPts];
      zz[c][1] = zst[(c * 64 + 2 * g + 1) * kTcPts];
    }
  };
  auto process = [&](const float2 (&zc)[C][2], int g) {
    const int ocol = g * 4;
    float hb[C][4];
    if (!lc.flag) {
#pragma unroll
      for (int c = 0; c < C; ++c) tmem_ld4(lc.taddr + c * kTwW + ocol, hb[c]);
      tc::tmem_ld_wait();
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const float wl = lds_f32(lc.fp + (FW_WL + ocol + i) * 4);
#pragma unroll
        for (int c = 0; c < C; ++c) hb[c][i] = wl * ub[c];
      }
    }
#pragma unroll
    for (int i = 0; i < 4; i += 2) {
      P2 zz[C], hv[C], zv[C];
#pragma unroll
      for (int c = 0; c < C; ++c) { zz[c].v = zc[c][i >> 1]; hv[c] = mk2(hb[c][i], hb[c][i + 1]); }
      chain_bwd<N1, N2, PURE, AK, P2>(lc.act, ch, zz, hv, zv);
#pragma unroll
      for (int c = 0; c < C; ++c) { hb[c][i] = zv[c].v.x; hb[c][i + 1] = zv[c].v.y; }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) store_half(tile_of(lc.tP, c, ocol), 0u, lc.p, ocol & 63, hb[c], false);
  };
  load(zb, lc.g0);
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
#pragma unroll
    for (int c = 0; c < C; ++c) { za[c][0] = zb[c][0]; za[c][1] = zb[c][1]; }
    if (g + 1 < lc.g1) load(zb, g + 1);
    process(za, g);
  }
}

// layer 0 reverse: adjoints of H^0 (TMEM X) -> Zbar^0 tiles of the value + first-derivative channels in P
template <int N1, int N2, bool PURE, int AK>
__device__ __forceinline__ void tw_l0_bwd_store_loop(const LoopW lc, const PassInfo<N1, N2> pi, const float* xp) {
  constexpr int C = 1 + N1 + N2;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = xp[k];
#pragma unroll 1
  for (int g = lc.g0; g < lc.g1; ++g) {
    const int col = g * 2;
    float hb[C][2];
#pragma unroll
    for (int c = 0; c < C; ++c) tmem_ld2(lc.taddr + c * kTwW + col, hb[c]);
    tc::tmem_ld_wait();
    float za[C], zb2[C];
    first_layer_elem_w<N1, N2>(lc.fp, pi, x, col, za);
    first_layer_elem_w<N1, N2>(lc.fp, pi, x, col + 1, zb2);
    P2 zz[C], hv[C], zv[C];
#pragma unroll
    for (int c = 0; c < C; ++c) { zz[c] = mk2(za[c], zb2[c]); hv[c] = mk2(hb[c][0], hb[c][1]); }
    chain_bwd<N1, N2, PURE, AK, P2>(lc.act, pi.ch, zz, hv, zv);
#pragma unroll
    for (int c = 0; c <= N1; ++c) {
      const float o2[2] = {zv[c].v.x, zv[c].v.y};
      store_half(tile_of(lc.tP, c, col), 0u, lc.p, col & 63, o2, false);
    }
  }
}

// ---- issuing-lane helpers (one elected lane of warp 0; phases live in shared memory) --------------------------------
__device__ __forceinline__ void tw_load(TwShared* cs, int b, uint32_t dst, const uint8_t* src, int n_tiles) {
  tc::mbar_arrive_expect_tx(&cs->bar_ld[b], (uint32_t)n_tiles * TB);
  for (int i = 0; i < n_tiles; ++i)
    tc::bulk_load_u(dst + (uint32_t)i * TB, src + (size_t)i * TB, TB, tc::smem_u32(&cs->bar_ld[b]));
}
__device__ __forceinline__ void tw_wait_ld(TwShared* cs, int b) {
  tc::mbar_wait(&cs->bar_ld[b], cs->ph_ld[b]);
  cs->ph_ld[b] ^= 1u;
}
__device__ __forceinline__ void tw_wait_free(TwShared* cs, int b) {
  tc::mbar_wait(&cs->bar_free[b], cs->ph_free[b]);
  cs->ph_free[b] ^= 1u;
}

// ---------------------------------------------------------------------------------------------------------------------
// forward of one network for the current tile.  Returns the updated parity of the MMA barrier.
template <int N1, int N2, bool PURE, int AK>
__device__ __noinline__ uint32_t tw_net_forward(TwShared* cs, const DevProblem* Pp, const DevTerm* tmp, int slot,
                                                int want_grad, uint32_t mma_phase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int C = 1 + N1 + N2;
  const DevTerm& tm = *tmp;
  const int net_id = tm.used_net[slot];
  const DevNet& net = reinterpret_cast<const DevNet*>(smem + cs->off_nets)[net_id];
  const DevChan& dc = tm.chan[slot];
  const float* fp = reinterpret_cast<const float*>(smem + cs->off_fp[net_id]);
  uint8_t* tP = smem + cs->off_P;
  uint8_t* tS = smem + cs->off_S;
  const Misc ms = misc_of(smem + cs->off_misc, cs->mx_dim, cs->mx_taps);
  const uint32_t tmem = cs->tmem;
  PassInfo<N1, N2> pi;
  load_pass<N1, N2>(pi, net, dc);
  const int TL = pi.TL;
  const Tid t = tid_of();
  const int tid = t.tid, hh = t.hh, p = t.p;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < pi.d_in) ? ms.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* hst = cs->hstash + (size_t)slot * (cs->tl_max + 1) * kTwMaxC * 2 * TB;
  float* zst = cs->zstash + (size_t)slot * cs->tl_max * kTwMaxC * 64 * kTcPts * 2;
  const uint8_t* wimg = cs->wpack + (size_t)cs->wimg[net_id] * kTwImgBytes;

  dbg_mark(cs, 10);
  float u[C];
#pragma unroll
  for (int c = 0; c < C; ++c) u[c] = 0.f;
  if (tid < C * kTcPts / 4) reinterpret_cast<float4*>(ms.scratch)[tid] = make_float4(0.f, 0.f, 0.f, 0.f);
  // first tensor layer's weights stream in behind the layer-0 epilogue (buffer S[1])
  if (tc::uni(t.warp) == 0) {
    const uint32_t u_S = tc::uni(tc::smem_u32(tS));
    const uint64_t u_w = tc::uni((uint64_t)wimg);
    const int u_nb = tc::uni((net.dims[1] + 63) >> 6);
    if (tc::elect_one()) {
      tc::fence_async_smem();
      tw_load(cs, 1, u_S + kTwImgBytes, (const uint8_t*)u_w, u_nb);
    }
    __syncwarp();
  }
  {
    const int ng = pi.n1w / 4;
    LoopW lc;
    lc.fp = tc::smem_u32(fp); lc.bt = lc.fp; lc.tP = tc::smem_u32(tP); lc.gb = nullptr;
    lc.taddr = tmem + t.lane_addr; lc.act = net.acts[0]; lc.p = p; lc.lane = t.lane;
    lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH); lc.flag = 0;
    tw_l0_fwd_loop<N1, N2, PURE, AK>(lc, pi, x);
  }
  for (int l = 1; l <= TL; ++l) {
    const int n_in = net.dims[l], n_out = net.dims[l + 1];
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    dbg_mark(cs, 11);
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
      const int u_nin = tc::uni(n_in), u_nout = tc::uni(n_out), u_wg = tc::uni(want_grad), u_l = tc::uni(l), u_TL = tc::uni(TL);
      const int u_nnext = tc::uni(l < TL ? net.dims[l + 1] : 0);
      const uint64_t u_hst = tc::uni((uint64_t)(hst + (size_t)(l - 1) * kTwMaxC * 2 * TB));
      const uint64_t u_w = tc::uni((uint64_t)wimg);
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const int nb_in = (u_nin + 63) >> 6;
        if (u_wg) {
#pragma unroll 1
          for (int c = 0; c < C; ++c)
            for (int kb = 0; kb < nb_in; ++kb)
              tc::bulk_store_u((void*)(u_hst + (uint64_t)(c * 2 + kb) * TB), u_P + (c * 2 + kb) * TB, TB);
          tc::bulk_commit();
        }
        if (u_l < u_TL)     // next layer's weights -> the other buffer (its last readers, layer l-1's MMAs, have retired)
          tw_load(cs, (u_l + 1) & 1, u_S + ((u_l + 1) & 1) * kTwImgBytes, (const uint8_t*)u_w + (size_t)u_l * kTwImgBytes,
                  (u_nnext + 63) >> 6);
        tw_wait_ld(cs, u_l & 1);
        const uint32_t idesc = tc::make_idesc(128, u_nout, 0, 0);
        const uint32_t wbuf = u_S + (u_l & 1) * kTwImgBytes;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const uint32_t d = u_tmem + c * kTwW;
#pragma unroll 1
          for (int kb = 0; kb < nb_in; ++kb) {
            const int nk = ((u_nin - kb * 64) < 64 ? (u_nin - kb * 64) : 64) >> 4;
            mma_chain(d, tc::make_desc(u_P + (c * 2 + kb) * TB, 0, 1024), tc::make_desc(wbuf + kb * TB, 0, 1024), 32, 32, nk,
                      idesc, kb > 0 ? 1u : 0u);
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
      if (tc::elect_one()) tc::bulk_wait_read0();       // stash copies have finished reading P
      __syncwarp();
    }
    __syncthreads();
    dbg_mark(cs, 14);
    const int ng = n_out / 4;
    LoopW lc;
    lc.fp = tc::smem_u32(fp); lc.bt = lc.fp + (FW_BT + (l - 1) * 128) * 4; lc.tP = tc::smem_u32(tP); lc.gb = nullptr;
    lc.taddr = tmem + t.lane_addr; lc.act = net.acts[l]; lc.p = p; lc.lane = t.lane;
    lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH); lc.flag = (l == TL) ? 1 : 0;
    float2* zl = want_grad ? reinterpret_cast<float2*>(zst + (size_t)(l - 1) * kTwMaxC * 64 * kTcPts * 2) + p : nullptr;
    tw_fwd_loop<N1, N2, PURE, AK>(lc, pi.ch, u, zl);
  }
  // ---- last layer (n -> 1, identity): combine the column parts of every point ---------------------------------
  if (want_grad && tm.n_used > 1) {
    // several passes share P: keep this pass's last hidden activations for its reverse sweep
    tc::fence_async_smem();
    __syncthreads();
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_P = tc::uni(tc::smem_u32(tP));
      const uint64_t u_hst = tc::uni((uint64_t)(hst + (size_t)TL * kTwMaxC * 2 * TB));
      const int u_nb = tc::uni((pi.nL + 63) >> 6);
      if (tc::elect_one()) {
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          for (int kb = 0; kb < u_nb; ++kb)
            tc::bulk_store_u((void*)(u_hst + (uint64_t)(c * 2 + kb) * TB), u_P + (c * 2 + kb) * TB, TB);
        tc::bulk_commit();
        tc::bulk_wait_read0();
      }
      __syncwarp();
    }
  }
  __syncthreads();
  dbg_mark(cs, 15);
#pragma unroll
  for (int c = 0; c < C; ++c) atomicAdd(&ms.scratch[c * kTcPts + p], u[c]);
  __syncthreads();
  if (hh == 0) {
#pragma unroll
    for (int c = 0; c < C; ++c) u[c] = ms.scratch[c * kTcPts + p];
    u[0] += fp[FW_BL];
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
  return mma_phase;
}

// reverse sweep of one network for the current tile (P still holds the last hidden activations)
template <int N1, int N2, bool PURE, int AK>
__device__ __noinline__ uint32_t tw_net_backward(TwShared* cs, const DevProblem* Pp, const DevTerm* tmp, int slot,
                                                 uint32_t mma_phase) {
  extern __shared__ __align__(1024) uint8_t smem[];
  constexpr int C = 1 + N1 + N2;
  constexpr uint32_t WG = (C <= 3) ? 384u : 0u;       // TMEM column of the weight-gradient accumulator
  constexpr uint32_t BC = (C <= 2) ? 256u : 128u;     // TMEM column of the bias-gradient column sums (16 columns)
  // dgrad channels whose TMEM columns hold the weight / bias gradient until it is flushed: issued after the flush
  constexpr uint32_t DEFER = (C == 4) ? 0x3u : ((C == 3) ? 0x2u : 0x0u);
  const DevTerm& tm = *tmp;
  const int net_id = tm.used_net[slot];
  const DevNet& net = reinterpret_cast<const DevNet*>(smem + cs->off_nets)[net_id];
  const DevChan& dc = tm.chan[slot];
  const float* fp = reinterpret_cast<const float*>(smem + cs->off_fp[net_id]);
  uint8_t* tP = smem + cs->off_P;
  uint8_t* tS = smem + cs->off_S;
  const Misc ms = misc_of(smem + cs->off_misc, cs->mx_dim, cs->mx_taps);
  const uint32_t tmem = cs->tmem;
  float* partial = cs->partial;
  PassInfo<N1, N2> pi;
  load_pass<N1, N2>(pi, net, dc);
  const int L = pi.L, TL = pi.TL;
  const Tid t = tid_of();
  const int tid = t.tid, hh = t.hh, p = t.p, lane = t.lane, q = t.q;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < pi.d_in) ? ms.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* hst = cs->hstash + (size_t)slot * (cs->tl_max + 1) * kTwMaxC * 2 * TB;
  const float* zst = cs->zstash + (size_t)slot * cs->tl_max * kTwMaxC * 64 * kTcPts * 2;
  const uint8_t* wimg = cs->wpack + (size_t)cs->wimg[net_id] * kTwImgBytes;

  dbg_mark(cs, 20);
  if (tm.n_used > 1) {
    // restore this pass's last hidden activations into P
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_P = tc::uni(tc::smem_u32(tP));
      const uint64_t u_hst = tc::uni((uint64_t)(hst + (size_t)TL * kTwMaxC * 2 * TB));
      const int u_nb = tc::uni((pi.nL + 63) >> 6);
      if (tc::elect_one()) {
        tc::fence_async_smem();
        tc::mbar_arrive_expect_tx(&cs->bar_ld[0], (uint32_t)(C * u_nb) * TB);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          for (int kb = 0; kb < u_nb; ++kb)
            tc::bulk_load_u(u_P + (c * 2 + kb) * TB, (const void*)(u_hst + (uint64_t)(c * 2 + kb) * TB), TB, tc::smem_u32(&cs->bar_ld[0]));
        tw_wait_ld(cs, 0);
      }
      __syncwarp();
    }
    __syncthreads();
  }
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
  // ---- last layer: bias gradient by warp sums; weight gradient  wbar_last[o] = sum_{c,p} ubar_c[p] H_c^{TL}[p][o]  on the
  // tensor core: D_c[o][0..15] = H_c^T U with U[p] = (hi, lo) bf16 pairs of ubar_0..ubar_(C-1) (columns 2c, 2c+1) ------------
  {
    float* gb_last = partial + net.b_off[L - 1];
    float* gw_last = partial + net.w_off[L - 1];
    if (hh == 0) {
      const float s = warp_sum<float>(ub[0]);
      if (lane == 0) atomicAdd(gb_last, s);
      uint32_t w[4];
#pragma unroll
      for (int c = 0; c < 4; ++c) w[c] = 0u;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        const uint32_t hi = tc::pack_bf16(ub[c], 0.f) & 0xffffu;
        const float r = ub[c] - __uint_as_float(hi << 16);
        w[c] = hi | (tc::pack_bf16(r, 0.f) << 16);
      }
      const uint32_t q0 = tc::smem_u32(tS);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + tc::swz_chunk(p, 0)), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + tc::swz_chunk(p, 1)), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
      const int u_nL = tc::uni(pi.nL);
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc(128, 16, 1, 1);
        const uint32_t a_lbo = (u_nL > 64) ? TB : 0u;
        const uint64_t db = tc::make_desc(u_S, 0, 1024);
#pragma unroll 1
        for (int c = 0; c < C; ++c)
          mma_chain(u_tmem + 16 * c, tc::make_desc(u_P + c * 2 * TB, a_lbo, 1024), db, 2048, 2048, kTcPts / 16, idesc, 0);
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    if (hh == 0) {
      const int o = q * 32 + lane;
      float acc = 0.f;
#pragma unroll
      for (int c = 0; c < C; ++c) {
        float v[2];
        tmem_ld2(tmem + t.lane_addr + 16 * c + 2 * c, v);
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
    float* gb = partial + net.b_off[l];
    float* gw = partial + net.w_off[l];
    dbg_mark(cs, 21);
    // this layer's input tiles of channels 0 and 1 stream into S0 / S1 behind the epilogue
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_S = tc::uni(tc::smem_u32(tS));
      const uint64_t u_hst = tc::uni((uint64_t)(hst + (size_t)(l - 1) * kTwMaxC * 2 * TB));
      const int u_nb = tc::uni((n_in + 63) >> 6);
      if (tc::elect_one()) {
        tc::fence_async_smem();
        tw_load(cs, 0, u_S, (const uint8_t*)u_hst, u_nb);
        if (C > 1) tw_load(cs, 1, u_S + kTwImgBytes, (const uint8_t*)u_hst + 2 * TB, u_nb);
      }
      __syncwarp();
    }
    {
      const int ng = n_out / 4;
      LoopW lc;
      lc.fp = tc::smem_u32(fp); lc.bt = lc.fp; lc.tP = tc::smem_u32(tP); lc.gb = gb;
      lc.taddr = tmem + t.lane_addr; lc.act = net.acts[l]; lc.p = p; lc.lane = lane;
      lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH); lc.flag = (l == TL) ? 1 : 0;
      const float2* zl = reinterpret_cast<const float2*>(zst + (size_t)(l - 1) * kTwMaxC * 64 * kTcPts * 2) + p;
      tw_bwd_loop<N1, N2, PURE, AK>(lc, pi.ch, ub, zl);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    dbg_mark(cs, 26);
    // wgrad: Wbar_l[o][k] = sum_c sum_p Zbar_c[p][o] H_c[p][k]  -> TMEM columns WG .. WG + n_in (lane = o)
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
      const int u_nin = tc::uni(n_in), u_nout = tc::uni(n_out), u_l = tc::uni(l);
      const uint32_t u_ones = tc::uni(tc::smem_u32(smem + cs->off_ones));
      const uint64_t u_hst = tc::uni((uint64_t)(hst + (size_t)(l - 1) * kTwMaxC * 2 * TB));
      const uint64_t u_w = tc::uni((uint64_t)wimg);
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const int nb_in = (u_nin + 63) >> 6;
        const uint32_t iwg = tc::make_idesc(128, u_nin, 1, 1);
        const uint32_t a_lbo = (u_nout > 64) ? TB : 0u;      // rows >= 64 of M: next tile, or alias rows - 64 when n_out <= 64
        if (C == 1) tw_load(cs, 1, u_S + kTwImgBytes, (const uint8_t*)u_w + (size_t)(u_l - 1) * kTwImgBytes, nb_in);
        int pend0 = 0, pend1 = 0;
#pragma unroll 1
        for (int c = 0; c < C; ++c) {
          const int b = c & 1;
          tw_wait_ld(cs, b);
          mma_chain(u_tmem + WG, tc::make_desc(u_P + (c * 2) * TB, a_lbo, 1024), tc::make_desc(u_S + b * kTwImgBytes, TB, 1024),
                    2048, 2048, kTcPts / 16, iwg, c > 0 ? 1u : 0u);
          tc::mma_commit(&cs->bar_free[b]);
          if (b) pend1 = 1; else pend0 = 1;
          if (c + 2 < C) {
            tw_wait_free(cs, b);
            if (b) pend1 = 0; else pend0 = 0;
            tw_load(cs, b, u_S + b * kTwImgBytes, (const uint8_t*)u_hst + (size_t)(c + 2) * 2 * TB, nb_in);
          } else if (c == C - 2) {
            // W_l for dgrad goes into the buffer wgrad releases first
            tw_wait_free(cs, b);
            if (b) pend1 = 0; else pend0 = 0;
            tw_load(cs, b, u_S + b * kTwImgBytes, (const uint8_t*)u_w + (size_t)(u_l - 1) * kTwImgBytes, nb_in);
          }
        }
        // bias gradient: bbar_l[o] = sum_p Zbar_0[p][o]  (B = the constant ones atom: SBO = 0, no k advance)
        mma_chain(u_tmem + BC, tc::make_desc(u_P, a_lbo, 1024), tc::make_desc(u_ones, 0, 0), 2048, 0, kTcPts / 16,
                  tc::make_idesc(128, 16, 1, 1), 0);
        if (pend0) tw_wait_free(cs, 0);
        if (pend1) tw_wait_free(cs, 1);
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    dbg_mark(cs, 27);
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    dbg_mark(cs, 28);
    // dgrad: Hbar_c[p][k] = sum_o Zbar_c[p][o] W_l[o][k] -> X (channel c at column c*128); W_l sits in S[C & 1]
    const int wb = C & 1;
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
      const int u_nin = tc::uni(n_in), u_nout = tc::uni(n_out);
      if (tc::elect_one()) {
        tw_wait_ld(cs, wb);
        const int nb_out = (u_nout + 63) >> 6;
        const uint32_t idg = tc::make_idesc(128, u_nin, 0, 1);
        const uint32_t wbuf = u_S + wb * kTwImgBytes;
#pragma unroll 1
        for (int c = C - 1; c >= 0; --c) {
          if ((DEFER >> c) & 1u) continue;
#pragma unroll 1
          for (int ob = 0; ob < nb_out; ++ob) {
            const int nk = ((u_nout - ob * 64) < 64 ? (u_nout - ob * 64) : 64) >> 4;
            mma_chain(u_tmem + c * kTwW, tc::make_desc(u_P + (c * 2 + ob) * TB, 0, 1024), tc::make_desc(wbuf + ob * 8192, TB, 1024),
                      32, 2048, nk, idg, ob > 0 ? 1u : 0u);
          }
        }
        if (DEFER == 0) tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    // flush the weight-gradient accumulator: TMEM lane = output neuron o, column = input neuron k
    {
      const int o = q * 32 + lane;
      if (hh == 0) {
        float v[2];
        tmem_ld2(tmem + t.lane_addr + BC, v);
        tc::tmem_ld_wait();
        if (o < n_out) atomicAdd(gb + o, v[0]);
      }
      const int part = n_in / kNH;
#pragma unroll 1
      for (int k0 = hh * part; k0 < (hh + 1) * part; k0 += 4) {
        float v[4];
        tmem_ld4(tmem + t.lane_addr + WG + k0, v);
        tc::tmem_ld_wait();
        if (o < n_out) {
#pragma unroll
          for (int i = 0; i < 4; ++i) atomicAdd(gw + o + (long long)n_out * (k0 + i), v[i]);
        }
      }
    }
    if (DEFER != 0) {
      // the remaining channels' adjoints land on the columns the weight / bias gradient just left
      tc::tc_fence_before();
      __syncthreads();
      if (tc::uni(t.warp) == 0) {
        const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
        const int u_nin = tc::uni(n_in), u_nout = tc::uni(n_out);
        if (tc::elect_one()) {
          tc::tc_fence_after();
          const int nb_out = (u_nout + 63) >> 6;
          const uint32_t idg = tc::make_idesc(128, u_nin, 0, 1);
          const uint32_t wbuf = u_S + wb * kTwImgBytes;
#pragma unroll 1
          for (int c = C - 1; c >= 0; --c) {
            if (!((DEFER >> c) & 1u)) continue;
#pragma unroll 1
            for (int ob = 0; ob < nb_out; ++ob) {
              const int nk = ((u_nout - ob * 64) < 64 ? (u_nout - ob * 64) : 64) >> 4;
              mma_chain(u_tmem + c * kTwW, tc::make_desc(u_P + (c * 2 + ob) * TB, 0, 1024),
                        tc::make_desc(wbuf + ob * 8192, TB, 1024), 32, 2048, nk, idg, ob > 0 ? 1u : 0u);
            }
          }
          tc::mma_commit(ms.bar_mma);
        }
        __syncwarp();
      }
    }
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    dbg_mark(cs, 29);
  }

  // ---- layer 0 reverse: Zbar^0 tiles, then  D[o][0..15] = Zbar_0^T [x | 1] + sum_j Zbar_(1+j)^T E_(dir1[j]) -----------------
  {
    float* gb0 = partial + net.b_off[0];
    float* gw0 = partial + net.w_off[0];
    constexpr bool kLo = (2 + N1) <= 4;         // a spare tile for the bf16 residual of the coordinates
    if (tid < kTcPts) {
      uint32_t hi[4], lo[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        hi[k] = tc::pack_bf16(x[2 * k], x[2 * k + 1]);
        lo[k] = tc::pack_bf16(x[2 * k] - __uint_as_float(hi[k] << 16), x[2 * k + 1] - __uint_as_float(hi[k] & 0xffff0000u));
      }
      const uint32_t q0 = tc::smem_u32(tS);
      const uint32_t c0a = tc::swz_chunk(p, 0), c1a = tc::swz_chunk(p, 1);
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + c0a), "r"(hi[0]), "r"(hi[1]), "r"(hi[2]), "r"(hi[3]) : "memory");
      asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + c1a), "r"(0x00003f80u), "r"(0u), "r"(0u), "r"(0u) : "memory");
      if (kLo) {
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + (1 + N1) * TB + c0a), "r"(lo[0]), "r"(lo[1]), "r"(lo[2]), "r"(lo[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(q0 + (1 + N1) * TB + c1a), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
      }
#pragma unroll
      for (int j = 0; j < N1; ++j) {
        const int d = pi.dir1[j];
        uint32_t w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int k = 0; k < 4; ++k) w[k] = (d == 2 * k) ? 0x00003f80u : ((d == 2 * k + 1) ? 0x3f800000u : 0u);
        const uint32_t qb = q0 + (1 + j) * TB;
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(qb + c0a), "r"(w[0]), "r"(w[1]), "r"(w[2]), "r"(w[3]) : "memory");
        asm volatile("st.shared.v4.b32 [%0], {%1,%2,%3,%4};" ::"r"(qb + c1a), "r"(0u), "r"(0u), "r"(0u), "r"(0u) : "memory");
      }
    }
    {
      const int ng = pi.n1w / 2;
      LoopW lc;
      lc.fp = tc::smem_u32(fp); lc.bt = lc.fp; lc.tP = tc::smem_u32(tP); lc.gb = gb0;
      lc.taddr = tmem + t.lane_addr; lc.act = net.acts[0]; lc.p = p; lc.lane = lane;
      lc.g0 = hh * (ng / kNH); lc.g1 = (hh + 1) * (ng / kNH); lc.flag = 0;
      tw_l0_bwd_store_loop<N1, N2, PURE, AK>(lc, pi, x);
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tc::uni(t.warp) == 0) {
      const uint32_t u_tmem = tc::uni(tmem), u_P = tc::uni(tc::smem_u32(tP)), u_S = tc::uni(tc::smem_u32(tS));
      const int u_n1w = tc::uni(pi.n1w);
      if (tc::elect_one()) {
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc(128, 16, 1, 1);
        const uint32_t a_lbo = (u_n1w > 64) ? TB : 0u;
        const uint64_t a0 = tc::make_desc(u_P, a_lbo, 1024);
        mma_chain(u_tmem + WG, a0, tc::make_desc(u_S, 0, 1024), 2048, 2048, kTcPts / 16, idesc, 0);
        if (kLo) mma_chain(u_tmem + WG, a0, tc::make_desc(u_S + (1 + N1) * TB, 0, 1024), 2048, 2048, kTcPts / 16, idesc, 1);
#pragma unroll 1
        for (int j = 0; j < N1; ++j)
          mma_chain(u_tmem + WG, tc::make_desc(u_P + (1 + j) * 2 * TB, a_lbo, 1024), tc::make_desc(u_S + (1 + j) * TB, 0, 1024),
                    2048, 2048, kTcPts / 16, idesc, 1);
        tc::mma_commit(ms.bar_mma);
      }
      __syncwarp();
    }
    wait_bar(ms.bar_mma, mma_phase);
    tc::tc_fence_after();
    if (hh == 0) {
      const int o = q * 32 + lane;
      float v[16];
      tc::tmem_ld16(tmem + t.lane_addr + WG, v);
      tc::tmem_ld_wait();
      if (o < pi.n1w) {
#pragma unroll
        for (int k = 0; k < PINN_MAX_IN; ++k)
          if (k < pi.d_in) atomicAdd(gw0 + o + (long long)pi.n1w * k, v[k]);
        atomicAdd(gb0 + o, v[8]);
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  dbg_mark(cs, 30);
  return mma_phase;
}


// channel structures the wide path instantiates (C <= 4)
#define PINN_TW_DISPATCH(n1, n2, pure, ak, CALL)                              \
  do {                                                                        \
    const int _ak = (ak);                                                     \
    const int _key = ((n1) * 8 + (n2)) * 2 + ((pure) ? 1 : 0);                \
    switch (_key) {                                                           \
      case (0 * 8 + 0) * 2: case (0 * 8 + 0) * 2 + 1: PINN_TC_CASE(0, 0, true, CALL);   \
      case (1 * 8 + 0) * 2: case (1 * 8 + 0) * 2 + 1: PINN_TC_CASE(1, 0, true, CALL);   \
      case (2 * 8 + 0) * 2: case (2 * 8 + 0) * 2 + 1: PINN_TC_CASE(2, 0, true, CALL);   \
      case (3 * 8 + 0) * 2: case (3 * 8 + 0) * 2 + 1: PINN_TC_CASE(3, 0, true, CALL);   \
      case (1 * 8 + 1) * 2: case (1 * 8 + 1) * 2 + 1: PINN_TC_CASE(1, 1, true, CALL);   \
      case (2 * 8 + 1) * 2 + 1: PINN_TC_CASE(2, 1, true, CALL);               \
      case (2 * 8 + 1) * 2: PINN_TC_CASE(2, 1, false, CALL);                  \
      default: break;                                                         \
    }                                                                         \
  } while (0)

// ---- weight packing: theta (fp32, out x in column-major) -> bf16 swizzled images [kb][128 rows o][64 k] ---------------
__global__ void __launch_bounds__(256) tw_pack_kernel(const TwPackArgs a) {
  const int img = blockIdx.x >> 3;
  const int idx = (blockIdx.x & 7) * 256 + threadIdx.x;     // 2048 16-byte chunks per image
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.tile_counter = a.counter_init;
  if (img >= a.n_images) return;
  const DevNet& net = a.prob->nets[a.img_net[img]];
  const int l = a.img_layer[img];
  const int n_in = net.dims[l], n_out = net.dims[l + 1];
  const long long woff = net.w_off[l];
  const int kb = idx >> 10, o = (idx >> 3) & 127, kc = idx & 7;
  float w[8];
#pragma unroll
  for (int e = 0; e < 8; ++e) {
    const int k = kb * 64 + kc * 8 + e;
    w[e] = (o < n_out && k < n_in) ? __ldg(&a.theta[woff + o + (long long)n_out * k]) : 0.f;
  }
  uint4 h;
  h.x = tc::pack_bf16(w[0], w[1]); h.y = tc::pack_bf16(w[2], w[3]);
  h.z = tc::pack_bf16(w[4], w[5]); h.w = tc::pack_bf16(w[6], w[7]);
  *reinterpret_cast<uint4*>(a.wpack + (size_t)img * kTwImgBytes + (size_t)kb * TB + tc::swz_chunk(o, kc)) = h;
}

__global__ void __launch_bounds__(kTcThreads, 1) tw_loss_grad_kernel(const __grid_constant__ TwArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ TwShared cs;
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
  if (tid == 0) {
    tc::mbar_init(ms.bar_mma, 1);
    tc::mbar_init(ms.bar_ld, 1);            // collocation-tile bulk loads (own barrier: the weight stream uses cs.bar_ld[])
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&cs.bar_ld[b], 1);
      tc::mbar_init(&cs.bar_free[b], 1);
      cs.ph_ld[b] = 0; cs.ph_free[b] = 0;
    }
    tc::fence_barrier_init();
    cs.tl_max = args.tl_max; cs.off_P = args.off_P; cs.off_S = args.off_S; cs.off_misc = args.off_misc; cs.off_ones = args.off_ones; cs.off_nets = args.off_nets; cs.mx_dim = args.mx_dim; cs.mx_taps = args.mx_taps;
    cs.partial = partial;
    cs.hstash = args.hstash + (long long)blockIdx.x * args.hstash_per_cta;
    cs.zstash = args.zstash + (long long)blockIdx.x * args.zstash_per_cta;
    cs.theta = theta; cs.wpack = args.wpack;
#ifdef PINN_DEBUG
    cs.dbg = (blockIdx.x == 0) ? args.dbg : nullptr;
#else
    cs.dbg = nullptr;
#endif
    cs.dbg_n = 0;
    for (int k = 0; k < PINN_MAX_NETS; ++k) { cs.off_fp[k] = args.off_fp[k]; cs.wimg[k] = args.wimg[k]; }
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
  {      // network descriptors: every layer of every sweep reads widths / offsets / activations
    const int nw = P.n_nets * (int)(sizeof(DevNet) / 4);
    const int* src = reinterpret_cast<const int*>(&P.nets[0]);
    int* dst = reinterpret_cast<int*>(smem + args.off_nets);
    for (int i = tid; i < nw; i += kTcThreads) dst[i] = __ldg(src + i);
  }
  if (tid < 64) {      // ones atom: row r (128 B) holds bf16 1.0 in logical column 0 = 16-byte chunk (0 ^ r)
    const int r = tid >> 3, ch = tid & 7;
    *reinterpret_cast<uint4*>(smem + args.off_ones + r * 128 + ch * 16) = make_uint4(ch == r ? 0x00003f80u : 0u, 0u, 0u, 0u);
  }
  // fp32 blocks of the first / last layers and the tensor-layer biases
  for (int kn = 0; kn < P.n_nets; ++kn) {
    if (args.off_fp[kn] < 0) continue;
    const DevNet& net = P.nets[kn];
    float* fp = reinterpret_cast<float*>(smem + args.off_fp[kn]);
    const int L = net.n_layers;
    for (int i = tid; i < FW_SIZE; i += kTcThreads) fp[i] = 0.f;
    __syncthreads();
    const int n1w = net.dims[1], d_in = net.dims[0];
    const long long w0 = net.w_off[0], b0 = net.b_off[0];
    for (int i = tid; i < n1w * d_in; i += kTcThreads) {
      const int o = i % n1w, k = i / n1w;
      fp[FW_W1 + o * 8 + k] = __ldg(&theta[w0 + i]);
    }
    for (int i = tid; i < n1w; i += kTcThreads) fp[FW_B1 + i] = __ldg(&theta[b0 + i]);
    for (int i = tid; i < (L - 2) * 128; i += kTcThreads) {
      const int l = 1 + i / 128, o = i & 127;
      if (o < net.dims[l + 1]) fp[FW_BT + (l - 1) * 128 + o] = __ldg(&theta[net.b_off[l] + o]);
    }
    const int nL = net.dims[L - 1];
    const long long wl = net.w_off[L - 1], bl = net.b_off[L - 1];
    for (int i = tid; i < nL; i += kTcThreads) fp[FW_WL + i] = __ldg(&theta[wl + i]);
    if (tid == 0) fp[FW_BL] = __ldg(&theta[bl]);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  if (tid == 0) cs.tmem = *ms.tmem_slot;
  __syncthreads();
  dbg_mark(&cs, 2);
  uint32_t phase = 0;
  uint32_t tile_ld_phase = 0;      // parity of the collocation-tile barrier (ms.bar_ld)

  // tiles are claimed dynamically after the first one (heavy PDE tiles come first in the enumeration, cheap boundary
  // tiles last): a static round-robin leaves the CTAs that drew an extra PDE tile 20 % behind the rest
  for (int tile = args.tile_begin + blockIdx.x; tile < args.tile_end;) {
    int ti = 0;
    while (ti + 1 < P.n_terms && tile >= args.dyn[ti + 1].tile0) ++ti;
    const DevTerm* tmp = &P.terms[ti];
    const DevTerm& tm = *tmp;
    const long long p0 = (long long)(tile - args.dyn[ti].tile0) * kTcPts;
    const long long n_pts = args.dyn[ti].n;
    const float* pts = reinterpret_cast<const float*>(args.dyn[ti].pts);
    const float* qw = reinterpret_cast<const float*>(args.dyn[ti].qw);
    if (tid < (int)((sizeof(DevTerm) + 127) / 128)) tc::prefetch_l1(reinterpret_cast<const char*>(tmp) + tid * 128);
    const int dim = tm.dim, n_taps = tm.n_taps, n_used = tm.n_used, weighted = tm.weighted;
    // collocation tile = one contiguous block of dim x 512 bytes: one bulk transfer of the TMA unit into the scratch array,
    // transposed to [row][point] by 128 threads (see tc_kernel.cu); partial / unaligned tiles take the per-element path
    const float* tile_src = pts + p0 * dim;
    const bool bulk_tile = (p0 + kTcPts <= n_pts) && dim <= kTwMaxC && ((reinterpret_cast<uintptr_t>(tile_src) & 15) == 0);
    if (bulk_tile) {
      if (tid == 0) {
        tc::mbar_arrive_expect_tx(ms.bar_ld, (uint32_t)(dim * kTcPts * 4));
        tc::bulk_load(ms.scratch, tile_src, (uint32_t)(dim * kTcPts * 4), ms.bar_ld);
      }
      wait_bar(ms.bar_ld, tile_ld_phase);
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
      PINN_TW_DISPATCH(k1, k2, pu, ak, (phase = tw_net_forward<A1, A2, PU, AK>(&cs, Pp, tmp, slot, want_grad ? 1 : 0, phase)));
    }

    dbg_mark(&cs, 4);
    // S0 / S1 are idle between the sweeps: stage the program text and the per-point value / adjoint arrays there
    const int n_instr = tm.n_instr;
    DevInstr* sprog = reinterpret_cast<DevInstr*>(smem + args.off_S);
    float* sval = reinterpret_cast<float*>(smem + args.off_S + 8192);
    const bool prog_sm = (size_t)8192 + (size_t)2 * n_instr * kTcPts * 4 <= (size_t)(2 * kTwImgBytes);
    if (prog_sm) {
      const int nw = n_instr * (int)(sizeof(DevInstr) / 4);
      const int* src = reinterpret_cast<const int*>(tm.prog);
      for (int i = tid; i < nw; i += kTcThreads) reinterpret_cast<int*>(sprog)[i] = __ldg(src + i);
      __syncthreads();
    }
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
      if (tid == 0) tc::bulk_wait0();             // operand-tile stash writes of this tile are complete before reloads
      __threadfence_block();
      __syncthreads();
      dbg_mark(&cs, 6);
      for (int slot = n_used - 1; slot >= 0; --slot) {
        const int k1 = tm.chan[slot].n1, k2 = tm.chan[slot].n2, pu = tm.chan[slot].pure;
        const int ak = args.net_ak[tm.used_net[slot]];
        PINN_TW_DISPATCH(k1, k2, pu, ak, (phase = tw_net_backward<A1, A2, PU, AK>(&cs, Pp, tmp, slot, phase)));
      }
    }
    // claim the next tile only now: claiming a tile ahead would hand the last cheap tiles to CTAs that still owe a heavy one
    if (tid == 0) cs.next_tile = atomicAdd(args.tile_counter, 1);
    __syncthreads();
    tile = cs.next_tile;
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
cudaError_t tw_pack_launch(const TwPackArgs& a, cudaStream_t st) {
  tw_pack_kernel<<<(a.n_images > 0 ? a.n_images : 1) * 8, 256, 0, st>>>(a);
  return cudaGetLastError();
}

cudaError_t tw_launch(const TwArgs& a, int grid, size_t smem, cudaStream_t st) {
  static size_t granted[64] = {0};
  cudaError_t e = ensure_dynamic_smem(tw_loss_grad_kernel, smem, granted);
  if (e != cudaSuccess) return e;
  return launch_fused_kernel(tw_loss_grad_kernel, a, grid, kTcThreads, smem, st, a.tail.state != nullptr);
}

}  // namespace pinn
