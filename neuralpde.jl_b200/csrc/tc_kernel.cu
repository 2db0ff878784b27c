// tc_kernel.cu -- fused PINN loss+gradient kernel, tcgen05 tensor-core path (sm_100a).
//
// One CTA (256 threads) owns a tile of 128 collocation points; a point is a TMEM lane and a
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
//
// The first (d -> n) and last (n -> 1) layers are tiny and stay on the CUDA cores inside the
// same epilogues.  Arithmetic modes: PINN_MODE_TC_BF16 (one MMA per product) and
// PINN_MODE_TC_SPLIT (forward operands split into bf16 hi + lo, three MMAs per product,
// which restores ~fp32 accuracy of the loss; the reverse sweep uses the hi parts).
//
// Replaces the same reference functions as the FFMA path (see ffma_kernel.cuh).
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "tc_types.h"
#include "ffma_kernel.cuh"   // act_eval, run_program, warp_sum
#include "tc_prims.cuh"

namespace pinn {

// ---- small helpers ---------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ float pick(const float* v, int idx) {
  float r = v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
  return r;
}
template <int N>
__device__ __forceinline__ void add_at(float* v, int idx, float x) {
#pragma unroll
  for (int i = 0; i < N; ++i) v[i] += (idx == i) ? x : 0.f;
}

__device__ __forceinline__ float bf16_hi(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

// store 8 consecutive columns (one 16-byte chunk) of a row into a swizzled tile, hi and optionally lo
__device__ __forceinline__ void store_chunk(uint8_t* tile_hi, uint8_t* tile_lo, int row, int chunk, const float (&v)[8],
                                            bool split) {
  const uint32_t off = tc::swz_chunk(row, chunk);
  uint4 h;
  h.x = tc::pack_bf16(v[0], v[1]); h.y = tc::pack_bf16(v[2], v[3]);
  h.z = tc::pack_bf16(v[4], v[5]); h.w = tc::pack_bf16(v[6], v[7]);
  *reinterpret_cast<uint4*>(tile_hi + off) = h;
  if (split) {
    float r[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) r[i] = v[i] - bf16_hi(v[i]);
    uint4 l;
    l.x = tc::pack_bf16(r[0], r[1]); l.y = tc::pack_bf16(r[2], r[3]);
    l.z = tc::pack_bf16(r[4], r[5]); l.w = tc::pack_bf16(r[6], r[7]);
    *reinterpret_cast<uint4*>(tile_lo + off) = l;
  }
}

__device__ __forceinline__ float load_bf16(const uint8_t* tile, int row, int col) {
  const __nv_bfloat16 b = *reinterpret_cast<const __nv_bfloat16*>(tile + tc::swz_off(row, col));
  return __bfloat162float(b);
}

// channel bookkeeping of one (term, network): value + N1 first + N2 second derivative channels
template <int N1, int N2>
struct Chan {
  static constexpr int C = 1 + N1 + N2;
  int sa[N2 > 0 ? N2 : 1], sb[N2 > 0 ? N2 : 1];
};

// post-activation channels from pre-activation channels (z[0] value, z[1..N1], z[1+N1..])
template <int N1, int N2>
__device__ __forceinline__ void chain_fwd(int act, const Chan<N1, N2>& ch, const float* z, float* h) {
  constexpr int M1 = (N1 > 0) ? N1 : 1;
  float a, d1, d2, d3;
  act_eval<float>(act, z[0], a, d1, d2, d3);
  h[0] = a;
#pragma unroll
  for (int i = 0; i < N1; ++i) h[1 + i] = d1 * z[1 + i];
#pragma unroll
  for (int s = 0; s < N2; ++s) {
    const float za = pick<M1>(z + 1, ch.sa[s]), zb = pick<M1>(z + 1, ch.sb[s]);
    h[1 + N1 + s] = d1 * z[1 + N1 + s] + d2 * za * zb;
  }
}

// adjoints of pre-activations from adjoints of post-activations
template <int N1, int N2>
__device__ __forceinline__ void chain_bwd(int act, const Chan<N1, N2>& ch, const float* z, const float* hb, float* zb) {
  constexpr int M1 = (N1 > 0) ? N1 : 1;
  float a, d1, d2, d3;
  act_eval<float>(act, z[0], a, d1, d2, d3);
  float acc0 = d1 * hb[0];
#pragma unroll
  for (int i = 0; i < N1; ++i) {
    acc0 += d2 * z[1 + i] * hb[1 + i];
    zb[1 + i] = d1 * hb[1 + i];
  }
#pragma unroll
  for (int s = 0; s < N2; ++s) {
    const float za = pick<M1>(z + 1, ch.sa[s]), zbb = pick<M1>(z + 1, ch.sb[s]);
    const float g = hb[1 + N1 + s];
    acc0 += (d2 * z[1 + N1 + s] + d3 * za * zbb) * g;
    add_at<M1>(zb + 1, ch.sa[s], d2 * zbb * g);
    add_at<M1>(zb + 1, ch.sb[s], d2 * za * g);
    zb[1 + N1 + s] = d1 * g;
  }
  zb[0] = acc0;
}

// everything a tile phase needs, gathered once
struct TileCtx {
  uint8_t* smem;
  uint8_t *P, *Q;
  float *Xs, *taps, *tapbar, *qws, *rres, *scratch;
  double* tsum;
  uint64_t *bar_mma, *bar_ld;
  uint32_t tmem;
  uint32_t mma_phase, ld_phase;
  int tid, warp, lane, q, hh, p;
  uint32_t lane_addr;   // (q*32) << 16
};

__device__ __forceinline__ void wait_mma(TileCtx& cx) {
  tc::mbar_wait(cx.bar_mma, cx.mma_phase);
  cx.mma_phase ^= 1u;
  tc::tc_fence_after();
}

// first layer pre-activations of 8 neurons [o0, o0+8): z[c][i]
template <int N1, int N2>
__device__ __forceinline__ void first_layer_z(const float* fp, const DevChan& dc, const float* x /*[d]*/, int d_in, int o0,
                                              float (&z)[1 + N1 + N2][8]) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int o = o0 + i;
    float s = fp[FP_B1 + o];
#pragma unroll
    for (int k = 0; k < PINN_MAX_IN; ++k)
      if (k < d_in) s = fmaf(fp[FP_W1 + o * 8 + k], x[k], s);
    z[0][i] = s;
#pragma unroll
    for (int j = 0; j < N1; ++j) z[1 + j][i] = fp[FP_W1 + o * 8 + dc.dir1[j]];
#pragma unroll
    for (int j = 0; j < N2; ++j) z[1 + N1 + j][i] = 0.f;
  }
}

// ---------------------------------------------------------------------------------------------------
// forward + backward of one network for the current tile, channel structure <N1, N2>
template <int N1, int N2>
__device__ __noinline__ void net_forward(TileCtx& cx, const TcArgs& args, const DevProblem& P, const DevTerm& tm, int slot,
                                         bool want_grad, uint8_t* stash) {
  constexpr int C = 1 + N1 + N2;
  const int net_id = tm.used_net[slot];
  const DevNet& net = P.nets[net_id];
  const DevChan& dc = tm.chan[slot];
  const TcNetSmem& ns = args.nets[net_id];
  const float* fp = reinterpret_cast<const float*>(cx.smem + ns.fp);
  const int L = net.n_layers;
  const int d_in = net.dims[0];
  const int TL = L - 2;                         // tensor layers 1..L-2
  const bool split = args.split != 0;
  Chan<N1, N2> ch;
#pragma unroll
  for (int s = 0; s < N2; ++s) { ch.sa[s] = dc.s_a[s]; ch.sb[s] = dc.s_b[s]; }
  const int p = cx.p, hh = cx.hh, tid = cx.tid;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < d_in) ? cx.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* stash_slot = stash + (size_t)slot * args.tl_max * kTcMaxC * kTileBytes;

  {
    // =============================== FORWARD ==========================================================
    float u[C];                                   // last-layer partial dot products of this thread
#pragma unroll
    for (int c = 0; c < C; ++c) u[c] = 0.f;
    {
      // ---- layer 0 on the CUDA cores ------------------------------------------------------------------
      const int n1w = net.dims[1];
      const int nch = n1w / 8, c0 = hh * (nch / 2), c1 = (hh + 1) * (nch / 2);
      for (int j = c0; j < c1; ++j) {
        float z[C][8], h[C][8];
        first_layer_z<N1, N2>(fp, dc, x, d_in, j * 8, z);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float zz[C], hv[C];
#pragma unroll
          for (int c = 0; c < C; ++c) zz[c] = z[c][i];
          chain_fwd<N1, N2>(net.acts[0], ch, zz, hv);
#pragma unroll
          for (int c = 0; c < C; ++c) h[c][i] = hv[c];
          if (TL == 0) {
            const float wl = fp[FP_WL + j * 8 + i];
#pragma unroll
            for (int c = 0; c < C; ++c) u[c] = fmaf(wl, hv[c], u[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          store_chunk(cx.P + c * kTileBytes, cx.Q + c * kTileBytes, p, j, h[c], split);
      }
    }
    // ---- tensor layers -------------------------------------------------------------------------------------
    for (int l = 1; l <= TL; ++l) {
      const int n_in = net.dims[l], n_out = net.dims[l + 1];
      tc::fence_async_smem();
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        if (want_grad) {
          for (int c = 0; c < C; ++c)
            tc::bulk_store(stash_slot + (size_t)((l - 1) * kTcMaxC + c) * kTileBytes, cx.P + c * kTileBytes, kTileBytes);
          tc::bulk_commit();
        }
        const uint32_t idesc = tc::make_idesc(128, n_out, 0, 0);
        const uint32_t whi = tc::smem_u32(cx.smem + ns.w_hi[l - 1]), wlo = tc::smem_u32(cx.smem + ns.w_lo[l - 1]);
        for (int c = 0; c < C; ++c) {
          const uint32_t ahi = tc::smem_u32(cx.P + c * kTileBytes), alo = tc::smem_u32(cx.Q + c * kTileBytes);
          const uint32_t d = cx.tmem + TM_X + c * 64;
          uint32_t acc = 0;
          for (int k = 0; k < n_in / 16; ++k) {
            tc::mma_bf16(d, tc::make_desc(ahi + k * 32, 0, 1024), tc::make_desc(whi + k * 32, 0, 1024), idesc, acc);
            acc = 1;
          }
          if (split) {
            for (int k = 0; k < n_in / 16; ++k)
              tc::mma_bf16(d, tc::make_desc(ahi + k * 32, 0, 1024), tc::make_desc(wlo + k * 32, 0, 1024), idesc, 1);
            for (int k = 0; k < n_in / 16; ++k)
              tc::mma_bf16(d, tc::make_desc(alo + k * 32, 0, 1024), tc::make_desc(whi + k * 32, 0, 1024), idesc, 1);
          }
        }
        tc::mma_commit(cx.bar_mma);
      }
      wait_mma(cx);
      if (tid == 0 && want_grad) tc::bulk_wait_read0();   // stash copies have finished reading P
      __syncthreads();
      // ---- epilogue: TMEM -> bias + activation chain -> next operand tiles ------------------------------------
      const int nch = n_out / 8, c0 = hh * (nch / 2), c1 = (hh + 1) * (nch / 2);
      for (int j = c0; j < c1; ++j) {
        float z[C][8], h[C][8];
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ld8(cx.tmem + cx.lane_addr + TM_X + c * 64 + j * 8, z[c]);
        tc::tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float zz[C], hv[C];
          zz[0] = z[0][i] + fp[FP_BT + (l - 1) * 64 + j * 8 + i];
#pragma unroll
          for (int c = 1; c < C; ++c) zz[c] = z[c][i];
          chain_fwd<N1, N2>(net.acts[l], ch, zz, hv);
#pragma unroll
          for (int c = 0; c < C; ++c) h[c][i] = hv[c];
          if (l == TL) {
            const float wl = fp[FP_WL + j * 8 + i];
#pragma unroll
            for (int c = 0; c < C; ++c) u[c] = fmaf(wl, hv[c], u[c]);
          }
        }
#pragma unroll
        for (int c = 0; c < C; ++c)
          store_chunk(cx.P + c * kTileBytes, cx.Q + c * kTileBytes, p, j, h[c], split);
      }
    }
    // ---- last layer (n -> 1, identity): combine the two column halves --------------------------------------------
    if (hh == 1) {
#pragma unroll
      for (int c = 0; c < C; ++c) cx.scratch[c * kTcPts + p] = u[c];
    }
    __syncthreads();
    if (hh == 0) {
#pragma unroll
      for (int c = 0; c < C; ++c) u[c] += cx.scratch[c * kTcPts + p];
      u[0] += fp[FP_BL];
      for (int t = 0; t < tm.n_taps; ++t)
        if (tm.tap_slot[t] == slot) {
          float v = u[0];
#pragma unroll
          for (int c = 1; c < C; ++c) v = (tm.tap_ch[t] == c) ? u[c] : v;
          cx.taps[t * kTcPts + p] = v;
        }
    }
    __syncthreads();
  }
}

template <int N1, int N2>
__device__ __noinline__ void net_backward(TileCtx& cx, const TcArgs& args, const DevProblem& P, const DevTerm& tm, int slot,
                                          float* partial, uint8_t* stash) {
  constexpr int C = 1 + N1 + N2;
  const int net_id = tm.used_net[slot];
  const DevNet& net = P.nets[net_id];
  const DevChan& dc = tm.chan[slot];
  const TcNetSmem& ns = args.nets[net_id];
  const float* fp = reinterpret_cast<const float*>(cx.smem + ns.fp);
  const int L = net.n_layers;
  const int d_in = net.dims[0];
  const int TL = L - 2;                         // tensor layers 1..L-2
  const bool split = args.split != 0;
  Chan<N1, N2> ch;
#pragma unroll
  for (int s = 0; s < N2; ++s) { ch.sa[s] = dc.s_a[s]; ch.sb[s] = dc.s_b[s]; }
  const int p = cx.p, hh = cx.hh, tid = cx.tid;
  float x[PINN_MAX_IN];
#pragma unroll
  for (int k = 0; k < PINN_MAX_IN; ++k) x[k] = (k < d_in) ? cx.Xs[dc.rows[k] * kTcPts + p] : 0.f;
  uint8_t* stash_slot = stash + (size_t)slot * args.tl_max * kTcMaxC * kTileBytes;

  // =============================== BACKWARD ==============================================================
  // adjoint of the network outputs per channel (every thread of the point needs it)
  float ub[C];
#pragma unroll
  for (int c = 0; c < C; ++c) ub[c] = 0.f;
  for (int t = 0; t < tm.n_taps; ++t)
    if (tm.tap_slot[t] == slot) {
      const float g = cx.tapbar[t * kTcPts + p];
#pragma unroll
      for (int c = 0; c < C; ++c) ub[c] += (tm.tap_ch[t] == c) ? g : 0.f;
    }
  const int nL = net.dims[L - 1];               // width of the last hidden layer
  // ---- last layer: bias and weight gradient on the CUDA cores -------------------------------------------------------
  if (hh == 0) {
    const float s = warp_sum<float>(ub[0]);
    if (cx.lane == 0) atomicAdd(&partial[net.b_off[L - 1]], s);
#pragma unroll
    for (int c = 0; c < C; ++c) cx.scratch[c * kTcPts + p] = ub[c];
  }
  __syncthreads();
  {
    // thread <-> (neuron o = tid & 63, quarter of the points): P still holds H^{L-2}
    const int o = tid & 63, part = tid >> 6;
    if (o < nL) {
      float acc = 0.f;
      for (int pp = part * 32; pp < part * 32 + 32; ++pp) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
          float hv = load_bf16(cx.P + c * kTileBytes, pp, o);
          if (split) hv += load_bf16(cx.Q + c * kTileBytes, pp, o);
          acc = fmaf(cx.scratch[c * kTcPts + pp], hv, acc);
        }
      }
      atomicAdd(&partial[net.w_off[L - 1] + o], acc);
    }
  }
  __syncthreads();

  // ---- tensor layers, last to first ------------------------------------------------------------------------------------
  for (int l = TL; l >= 1; --l) {
    const int n_in = net.dims[l], n_out = net.dims[l + 1];
    const uint32_t whi = tc::smem_u32(cx.smem + ns.w_hi[l - 1]);
    if (tid == 0) {
      // reload this layer's input tiles H^{l-1} (bf16 hi) from the stash into Q
      tc::mbar_arrive_expect_tx(cx.bar_ld, C * kTileBytes);
      for (int c = 0; c < C; ++c)
        tc::bulk_load(cx.Q + c * kTileBytes, stash_slot + (size_t)((l - 1) * kTcMaxC + c) * kTileBytes, kTileBytes, cx.bar_ld);
    }
    tc::mbar_wait(cx.bar_ld, cx.ld_phase);
    cx.ld_phase ^= 1u;
    // recompute pre-activations in groups of <= 32 columns and turn output adjoints into Zbar tiles
    for (int g0 = 0; g0 < n_out; g0 += 32) {
      const int gw = (n_out - g0) < 32 ? (n_out - g0) : 32;     // 32 or 16
      tc::tc_fence_before();
      __syncthreads();
      if (tid == 0) {
        tc::tc_fence_after();
        const uint32_t idesc = tc::make_idesc(128, gw, 0, 0);
        for (int c = 0; c < C; ++c) {
          const uint32_t a = tc::smem_u32(cx.Q + c * kTileBytes);
          uint32_t acc = 0;
          for (int k = 0; k < n_in / 16; ++k) {
            tc::mma_bf16(cx.tmem + TM_Y + c * 32, tc::make_desc(a + k * 32, 0, 1024),
                         tc::make_desc(whi + g0 * 128 + k * 32, 0, 1024), idesc, acc);
            acc = 1;
          }
        }
        tc::mma_commit(cx.bar_mma);
      }
      wait_mma(cx);
      const int nch = gw / 8, c0 = hh * (nch / 2), c1 = (hh + 1) * (nch / 2);
      for (int jj = c0; jj < c1; ++jj) {
        const int ocol = g0 + jj * 8;
        float z[C][8], hb[C][8], zb[C][8];
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ld8(cx.tmem + cx.lane_addr + TM_Y + c * 32 + jj * 8, z[c]);
        if (l < TL) {
#pragma unroll
          for (int c = 0; c < C; ++c) tc::tmem_ld8(cx.tmem + cx.lane_addr + TM_X + c * 64 + ocol, hb[c]);
        }
        tc::tmem_ld_wait();
        if (l == TL) {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float wl = fp[FP_WL + ocol + i];
#pragma unroll
            for (int c = 0; c < C; ++c) hb[c][i] = wl * ub[c];
          }
        }
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float zz[C], hv[C], zv[C];
          zz[0] = z[0][i] + fp[FP_BT + (l - 1) * 64 + ocol + i];
#pragma unroll
          for (int c = 1; c < C; ++c) zz[c] = z[c][i];
#pragma unroll
          for (int c = 0; c < C; ++c) hv[c] = hb[c][i];
          chain_bwd<N1, N2>(net.acts[l], ch, zz, hv, zv);
#pragma unroll
          for (int c = 0; c < C; ++c) zb[c][i] = zv[c];
          const float bs = warp_sum<float>(zv[0]);
          if (cx.lane == 0) atomicAdd(&partial[net.b_off[l] + ocol + i], bs);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) store_chunk(cx.P + c * kTileBytes, cx.P, p, ocol >> 3, zb[c], false);
      }
    }
    tc::fence_async_smem();
    tc::tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc::tc_fence_after();
      // dgrad: Hbar_c = Zbar_c * W_l  -> X
      const uint32_t idg = tc::make_idesc(128, n_in, 0, 1);
      for (int c = 0; c < C; ++c) {
        const uint32_t a = tc::smem_u32(cx.P + c * kTileBytes);
        uint32_t acc = 0;
        for (int k = 0; k < n_out / 16; ++k) {
          tc::mma_bf16(cx.tmem + TM_X + c * 64, tc::make_desc(a + k * 32, 0, 1024),
                       tc::make_desc(whi + k * 2048, 0, 1024), idg, acc);
          acc = 1;
        }
      }
      // wgrad: Wbar_l = sum_c Zbar_c^T * H_c -> Y (rows >= 64 alias rows - 64 through LBO = 0)
      const uint32_t iwg = tc::make_idesc(128, n_in, 1, 1);
      uint32_t acc = 0;
      for (int c = 0; c < C; ++c) {
        const uint32_t a = tc::smem_u32(cx.P + c * kTileBytes), b = tc::smem_u32(cx.Q + c * kTileBytes);
        for (int k = 0; k < kTcPts / 16; ++k) {
          tc::mma_bf16(cx.tmem + TM_Y, tc::make_desc(a + k * 2048, 0, 1024), tc::make_desc(b + k * 2048, 0, 1024), iwg, acc);
          acc = 1;
        }
      }
      tc::mma_commit(cx.bar_mma);
    }
    wait_mma(cx);
    // flush the weight-gradient tile: TMEM lane = output neuron o, column = input neuron k
    if (cx.q < 2) {
      const int o = cx.q * 32 + cx.lane;
      const int half = n_in / 2;
      for (int k0 = hh * half; k0 < (hh + 1) * half; k0 += 8) {
        float v[8];
        tc::tmem_ld8(cx.tmem + cx.lane_addr + TM_Y + k0, v);
        tc::tmem_ld_wait();
        if (o < n_out) {
#pragma unroll
          for (int i = 0; i < 8; ++i) partial[net.w_off[l] + o + (long long)n_out * (k0 + i)] += v[i];
        }
      }
    }
  }

  // ---- layer 0 backward on the CUDA cores ---------------------------------------------------------------------------------------
  {
    tc::tc_fence_before();
    __syncthreads();
    tc::tc_fence_after();
    const int n1w = net.dims[1];
    const int nch = n1w / 8, c0 = hh * (nch / 2), c1 = (hh + 1) * (nch / 2);
    for (int j = c0; j < c1; ++j) {
      float z[C][8], hb[C][8];
      first_layer_z<N1, N2>(fp, dc, x, d_in, j * 8, z);
      if (TL > 0) {
#pragma unroll
        for (int c = 0; c < C; ++c) tc::tmem_ld8(cx.tmem + cx.lane_addr + TM_X + c * 64 + j * 8, hb[c]);
        tc::tmem_ld_wait();
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float wl = fp[FP_WL + j * 8 + i];
#pragma unroll
          for (int c = 0; c < C; ++c) hb[c][i] = wl * ub[c];
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int o = j * 8 + i;
        float zz[C], hv[C], zv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) { zz[c] = z[c][i]; hv[c] = hb[c][i]; }
        chain_bwd<N1, N2>(net.acts[0], ch, zz, hv, zv);
        // Wbar_0[o][k] = sum_p zbar_0 x_k + zbar_(channel of direction k);  bbar_0[o] = sum_p zbar_0
        const float bs = warp_sum<float>(zv[0]);
        if (cx.lane == 0) atomicAdd(&partial[net.b_off[0] + o], bs);
#pragma unroll
        for (int k = 0; k < PINN_MAX_IN; ++k) {
          if (k < d_in) {
            float g = zv[0] * x[k];
#pragma unroll
            for (int jd = 0; jd < N1; ++jd) g += (dc.dir1[jd] == k) ? zv[1 + jd] : 0.f;
            g = warp_sum<float>(g);
            if (cx.lane == 0) atomicAdd(&partial[net.w_off[0] + o + (long long)n1w * k], g);
          }
        }
      }
    }
  }
  __syncthreads();
}

#define PINN_TC_DISPATCH(n1, n2, CALL)                                        \
  do {                                                                        \
    const int _key = (n1) * 8 + (n2);                                         \
    switch (_key) {                                                           \
      case 0 * 8 + 0: { constexpr int A1 = 0, A2 = 0; CALL; } break;          \
      case 1 * 8 + 0: { constexpr int A1 = 1, A2 = 0; CALL; } break;          \
      case 2 * 8 + 0: { constexpr int A1 = 2, A2 = 0; CALL; } break;          \
      case 3 * 8 + 0: { constexpr int A1 = 3, A2 = 0; CALL; } break;          \
      case 4 * 8 + 0: { constexpr int A1 = 4, A2 = 0; CALL; } break;          \
      case 1 * 8 + 1: { constexpr int A1 = 1, A2 = 1; CALL; } break;          \
      case 2 * 8 + 1: { constexpr int A1 = 2, A2 = 1; CALL; } break;          \
      case 3 * 8 + 1: { constexpr int A1 = 3, A2 = 1; CALL; } break;          \
      case 2 * 8 + 2: { constexpr int A1 = 2, A2 = 2; CALL; } break;          \
      default: break;                                                         \
    }                                                                         \
  } while (0)

__global__ void __launch_bounds__(kTcThreads, 1) tc_loss_grad_kernel(const TcArgs args) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const int tid = threadIdx.x;
  const DevProblem& P = *args.prob;
  TileCtx cx;
  cx.smem = smem;
  cx.P = smem + args.off_P;
  cx.Q = smem + args.off_Q;
  uint8_t* misc = smem + args.off_misc;
  cx.Xs = reinterpret_cast<float*>(misc);             misc += PINN_MAX_DIM * kTcPts * 4;
  cx.taps = reinterpret_cast<float*>(misc);           misc += kTcMaxTaps * kTcPts * 4;
  cx.tapbar = reinterpret_cast<float*>(misc);         misc += kTcMaxTaps * kTcPts * 4;
  cx.scratch = reinterpret_cast<float*>(misc);        misc += kTcMaxC * kTcPts * 4;
  cx.qws = reinterpret_cast<float*>(misc);            misc += kTcPts * 4;
  cx.rres = reinterpret_cast<float*>(misc);           misc += kTcPts * 4;
  cx.tsum = reinterpret_cast<double*>(misc);          misc += PINN_MAX_TERMS * 8;
  cx.bar_mma = reinterpret_cast<uint64_t*>(misc);     misc += 8;
  cx.bar_ld = reinterpret_cast<uint64_t*>(misc);      misc += 8;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(misc);
  cx.tid = tid; cx.warp = tid >> 5; cx.lane = tid & 31; cx.q = cx.warp & 3; cx.hh = cx.warp >> 2;
  cx.p = cx.q * 32 + cx.lane;
  cx.lane_addr = (uint32_t)(cx.q * 32) << 16;
  cx.mma_phase = 0; cx.ld_phase = 0;

  float* partial = args.partial + (long long)blockIdx.x * P.n_theta;
  uint8_t* stash = args.stash + (long long)blockIdx.x * args.stash_per_cta;
  const bool want_grad = (args.mode == 0);
  const float* theta = args.theta;

  // ---- per-CTA setup --------------------------------------------------------------------------------------------------------
  if (tid == 0) {
    tc::mbar_init(cx.bar_mma, 1);
    tc::mbar_init(cx.bar_ld, 1);
    tc::fence_barrier_init();
  }
  if (cx.warp == 0) tc::tmem_alloc<512>(tmem_slot);
  if (want_grad)
    for (long long i = tid; i < P.n_theta; i += kTcThreads) partial[i] = 0.f;
  if (tid < PINN_MAX_TERMS) cx.tsum[tid] = 0.0;
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
    for (int i = tid; i < n1w * d_in; i += kTcThreads) {
      const int o = i % n1w, k = i / n1w;
      fp[FP_W1 + o * 8 + k] = __ldg(&theta[net.w_off[0] + i]);
    }
    for (int i = tid; i < n1w; i += kTcThreads) fp[FP_B1 + i] = __ldg(&theta[net.b_off[0] + i]);
    for (int l = 1; l <= L - 2; ++l) {
      const int n_in = net.dims[l], n_out = net.dims[l + 1];
      uint8_t* thi = smem + ns.w_hi[l - 1];
      uint8_t* tlo = smem + ns.w_lo[l - 1];
      for (int i = tid; i < 64 * 64; i += kTcThreads) {
        const int o = i & 63, k = i >> 6;
        float w = (o < n_out && k < n_in) ? __ldg(&theta[net.w_off[l] + o + (long long)n_out * k]) : 0.f;
        const __nv_bfloat16 h = __float2bfloat16_rn(w);
        *reinterpret_cast<__nv_bfloat16*>(thi + tc::swz_off(o, k)) = h;
        if (args.split) *reinterpret_cast<__nv_bfloat16*>(tlo + tc::swz_off(o, k)) = __float2bfloat16_rn(w - __bfloat162float(h));
      }
      for (int i = tid; i < n_out; i += kTcThreads) fp[FP_BT + (l - 1) * 64 + i] = __ldg(&theta[net.b_off[l] + i]);
    }
    const int nL = net.dims[L - 1];
    for (int i = tid; i < nL; i += kTcThreads) fp[FP_WL + i] = __ldg(&theta[net.w_off[L - 1] + i]);
    if (tid == 0) fp[FP_BL] = __ldg(&theta[net.b_off[L - 1]]);
  }
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  cx.tmem = *tmem_slot;

  for (int tile = args.tile_begin + blockIdx.x; tile < args.tile_end; tile += gridDim.x) {
    int ti = 0;
    while (ti + 1 < P.n_terms && tile >= args.dyn[ti + 1].tile0) ++ti;
    const DevTerm& tm = P.terms[ti];
    const long long p0 = (long long)(tile - args.dyn[ti].tile0) * kTcPts;
    const long long n_pts = args.dyn[ti].n;
    const float* pts = reinterpret_cast<const float*>(args.dyn[ti].pts);
    const float* qw = reinterpret_cast<const float*>(args.dyn[ti].qw);
    for (int i = tid; i < tm.dim * kTcPts; i += kTcThreads) {
      int pp = i / tm.dim, r = i - pp * tm.dim;
      long long gp = p0 + pp;
      if (gp >= n_pts) gp = n_pts - 1;
      cx.Xs[r * kTcPts + pp] = pts[gp * tm.dim + r];
    }
    if (tid < kTcPts) {
      long long gp = p0 + tid;
      float w = 0.f;
      if (gp < n_pts) w = tm.weighted ? qw[gp] : 1.f;
      cx.qws[tid] = w;
    }
    for (int i = tid; i < tm.n_taps * kTcPts; i += kTcThreads) cx.tapbar[i] = 0.f;
    __syncthreads();

    for (int slot = 0; slot < tm.n_used; ++slot) {
      const DevChan& dc = tm.chan[slot];
      PINN_TC_DISPATCH(dc.n1, dc.n2, (net_forward<A1, A2>(cx, args, P, tm, slot, want_grad, stash)));
    }

    // ---- residual program, loss partial, tap adjoints (threads 0..127: one point each) -----------------------------------------
    if (tid < kTcPts) {
      float pbar[PINN_MAX_PARAMS];
#pragma unroll
      for (int j = 0; j < PINN_MAX_PARAMS; ++j) pbar[j] = 0.f;
      const float r = run_program<float, kTcPts>(tm, theta + P.param_off, cx.Xs, cx.taps, cx.tapbar, pbar, tid, want_grad);
      const float w = cx.qws[tid];
      double s = (double)w * (double)r * (double)r;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (cx.lane == 0) atomicAdd(&cx.tsum[ti], s);
      if (args.mode == 2) {
        long long gp = p0 + tid;
        if (gp < n_pts) args.resid_out[gp] = r;
      }
      if (want_grad) {
        const float g = (float)args.seed[ti] * w * 2.f * r;
        for (int t = 0; t < tm.n_taps; ++t) cx.tapbar[t * kTcPts + tid] *= g;
        for (int j = 0; j < P.n_params; ++j) {
          float v = warp_sum<float>(pbar[j] * g);
          if (cx.lane == 0) atomicAdd(&partial[P.param_off + j], v);
        }
      }
    }
    __syncthreads();

    if (want_grad) {
      if (tid == 0) tc::bulk_wait0();             // stash writes of this tile are complete before reloads
      __syncthreads();
      for (int slot = tm.n_used - 1; slot >= 0; --slot) {
        const DevChan& dc = tm.chan[slot];
        if (tm.n_used > 1) {
          // P must hold this slot's last hidden activations again: recompute its forward
          PINN_TC_DISPATCH(dc.n1, dc.n2, (net_forward<A1, A2>(cx, args, P, tm, slot, false, stash)));
        }
        PINN_TC_DISPATCH(dc.n1, dc.n2, (net_backward<A1, A2>(cx, args, P, tm, slot, partial, stash)));
      }
    }
  }

  tc::tc_fence_before();
  __syncthreads();
  if (tid < PINN_MAX_TERMS) args.term_sums[(long long)blockIdx.x * PINN_MAX_TERMS + tid] = cx.tsum[tid];
  if (cx.warp == 0) tc::tmem_dealloc<512>(cx.tmem);
}

// ---- host side ------------------------------------------------------------------------------------------------------------------
size_t tc_misc_bytes() {
  return (size_t)PINN_MAX_DIM * kTcPts * 4 + 2 * (size_t)kTcMaxTaps * kTcPts * 4 + (size_t)kTcMaxC * kTcPts * 4 +
         2 * kTcPts * 4 + PINN_MAX_TERMS * 8 + 16 + 16;
}

cudaError_t tc_launch(const TcArgs& a, int grid, size_t smem, cudaStream_t st) {
  cudaError_t e = cudaFuncSetAttribute(tc_loss_grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  tc_loss_grad_kernel<<<grid, kTcThreads, smem, st>>>(a);
  return cudaGetLastError();
}

}  // namespace pinn
