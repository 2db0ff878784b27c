// tc_common.cuh -- device helpers shared by the tcgen05 kernels (tc_kernel.cu: widths <= 64, all operands resident;
// tc_wide_kernel.cu: 128-wide layers, streamed weights): packed fp32x2 arithmetic, the forward-mode tap chain rule
// and its adjoint, TMEM loads, swizzled-tile stores, warp reduce-scatter, MMA issue helpers, dispatch macro.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "tc_types.h"
#include "ffma_kernel.cuh"   // act_eval, run_program, warp_sum
#include "tc_prims.cuh"

namespace pinn {

// ---- small helpers ---------------------------------------------------------------------------------
constexpr int kNH = kTcThreads / 128;   // warps per TMEM lane quadrant: each takes 1/kNH of the columns
#ifndef PINN_TC_GW
#define PINN_TC_GW 4
#endif
constexpr int GW = PINN_TC_GW;          // columns per epilogue granule (4 or 2)
#ifndef PINN_TC_GWB
#define PINN_TC_GWB 2
#endif
constexpr int GWB = PINN_TC_GWB;        // granule of the tensor-layer reverse epilogue (register-heaviest loop)

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

// activation value and first three derivatives.  AK = 1: tanh, 2: sigmoid (branch-free, built on
// ex2.approx + rcp.approx: absolute error ~1e-7, far below the bf16-split operand noise);
// AK = 0: any activation through the accurate generic evaluator.
template <int AK>
__device__ __forceinline__ void act_eval_tc(int act, float z, float& a, float& d1, float& d2, float& d3) {
#ifdef PINN_EXP_NO_ACT
  a = z; d1 = 1.f; d2 = 0.5f; d3 = 0.25f; return;
#endif
  if (AK == 1) {
    const float e = __expf(2.f * z);
    const float t = 1.f - __fdividef(2.f, e + 1.f);
    const float s = fmaf(-t, t, 1.f);
    a = t; d1 = s; d2 = -2.f * t * s; d3 = s * fmaf(6.f * t, t, -2.f);
  } else if (AK == 2) {
    const float g = __fdividef(1.f, 1.f + __expf(-z));
    const float g1 = g * (1.f - g);
    a = g; d1 = g1; d2 = g1 * fmaf(-2.f, g, 1.f); d3 = g1 * fmaf(-6.f, g1, 1.f);
  } else {
    act_eval<float>(act, z, a, d1, d2, d3);
  }
}
__device__ __forceinline__ int act_kind(int act) { return act == PINN_ACT_TANH ? 1 : (act == PINN_ACT_SIGMOID ? 2 : 0); }

__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t x, uint32_t y) {
  asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(x), "r"(y) : "memory");
}

// store 4 consecutive columns (half of a 16-byte chunk) of a row into a swizzled tile; with `split`
// also the bf16 residual v - bf16(v) into the lo tile
__device__ __forceinline__ void store_half(uint32_t tile_hi, uint32_t tile_lo, int row, int col0, const float (&v)[4],
                                           bool split) {
#ifdef PINN_EXP_NO_STS
  asm volatile("" ::"f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]));
  return;
#endif
  const uint32_t off = tc::swz_chunk(row, col0 >> 3) + ((col0 & 4) << 1);
  const uint32_t hx = tc::pack_bf16(v[0], v[1]), hy = tc::pack_bf16(v[2], v[3]);
  sts_v2(tile_hi + off, hx, hy);
  if (split) {
    const uint32_t lx = tc::pack_bf16(v[0] - __uint_as_float(hx << 16), v[1] - __uint_as_float(hx & 0xffff0000u));
    const uint32_t ly = tc::pack_bf16(v[2] - __uint_as_float(hy << 16), v[3] - __uint_as_float(hy & 0xffff0000u));
    sts_v2(tile_lo + off, lx, ly);
  }
}

__device__ __forceinline__ void tmem_ld4(uint32_t taddr, float (&v)[4]) {
#ifdef PINN_EXP_NO_LDTM
  v[0] = __uint_as_float(taddr & 0x3fffffu) * 1e-9f; v[1] = v[0] + 1e-3f; v[2] = v[0] - 1e-3f; v[3] = v[0] * 0.5f;
  return;
#endif
  uint32_t r[4];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(taddr)
               : "memory");
#pragma unroll
  for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void tmem_ld2(uint32_t taddr, float (&v)[2]) {
  uint32_t r[2];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r[0]), "=r"(r[1]) : "r"(taddr) : "memory");
  v[0] = __uint_as_float(r[0]); v[1] = __uint_as_float(r[1]);
}
__device__ __forceinline__ void tmem_ldg(uint32_t taddr, float (&v)[4]) { tmem_ld4(taddr, v); }
__device__ __forceinline__ void tmem_ldg(uint32_t taddr, float (&v)[2]) { tmem_ld2(taddr, v); }

// 2-column variant of store_half
__device__ __forceinline__ void store_half(uint32_t tile_hi, uint32_t tile_lo, int row, int col0, const float (&v)[2],
                                           bool split) {
  const uint32_t off = tc::swz_chunk(row, col0 >> 3) + ((col0 & 6) << 1);
  const uint32_t hx = tc::pack_bf16(v[0], v[1]);
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_hi + off), "r"(hx) : "memory");
  if (split) {
    const uint32_t lx = tc::pack_bf16(v[0] - __uint_as_float(hx << 16), v[1] - __uint_as_float(hx & 0xffff0000u));
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(tile_lo + off), "r"(lx) : "memory");
  }
}

// ---- packed fp32x2 arithmetic (Blackwell FFMA2 / FMUL2 / FADD2): two columns per instruction ---------------
struct P2 { float2 v; };
__device__ __forceinline__ P2 mk2(float a, float b) { P2 r; r.v = make_float2(a, b); return r; }
__device__ __forceinline__ P2 splat2(float a) { return mk2(a, a); }
__device__ __forceinline__ P2 operator*(P2 a, P2 b) { P2 r; r.v = __fmul2_rn(a.v, b.v); return r; }
__device__ __forceinline__ P2 operator+(P2 a, P2 b) { P2 r; r.v = __fadd2_rn(a.v, b.v); return r; }
__device__ __forceinline__ P2 vfma(P2 a, P2 b, P2 c) { P2 r; r.v = __ffma2_rn(a.v, b.v, c.v); return r; }
__device__ __forceinline__ float vfma(float a, float b, float c) { return fmaf(a, b, c); }
template <typename T> __device__ __forceinline__ T vsplat(float x);
template <> __device__ __forceinline__ float vsplat<float>(float x) { return x; }
template <> __device__ __forceinline__ P2 vsplat<P2>(float x) { return splat2(x); }
__device__ __forceinline__ float ex2_approx(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcp_approx(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// tanh and its first three derivatives for two columns at once
__device__ __forceinline__ void tanh_eval2(P2 z, P2& a, P2& d1, P2& d2, P2& d3) {
  const P2 zz = z * splat2(2.8853900817779268f);                 // 2 * log2(e)
  const P2 den = mk2(ex2_approx(zz.v.x), ex2_approx(zz.v.y)) + splat2(1.f);
  const P2 r = mk2(rcp_approx(den.v.x), rcp_approx(den.v.y));
  const P2 t = vfma(r, splat2(-2.f), splat2(1.f));
  const P2 s = vfma(t * splat2(-1.f), t, splat2(1.f));
  a = t; d1 = s;
  d2 = (t * s) * splat2(-2.f);
  d3 = s * vfma(t * splat2(6.f), t, splat2(-2.f));
}

// channel bookkeeping of one (term, network): value + N1 first + N2 second derivative channels.
// PURE: second-derivative channel s is d2/dx_s^2 of first-derivative channel s (no index selects).
template <int N1, int N2>
struct Chan {
  int sa[N2 > 0 ? N2 : 1], sb[N2 > 0 ? N2 : 1];
};

template <typename T, int N>
__device__ __forceinline__ T pickT(const T* v, int idx) {
  T r = v[0];
#pragma unroll
  for (int i = 1; i < N; ++i) r = (idx == i) ? v[i] : r;
  return r;
}

// activation dispatch on the value type: packed tanh for P2, scalar evaluators otherwise
template <int AK>
__device__ __forceinline__ void act_any(int act, float z, float& a, float& d1, float& d2, float& d3) {
  act_eval_tc<AK>(act, z, a, d1, d2, d3);
}
template <int AK>
__device__ __forceinline__ void act_any(int act, P2 z, P2& a, P2& d1, P2& d2, P2& d3) {
  if (AK == 1) {
    tanh_eval2(z, a, d1, d2, d3);
  } else {
    float ax, d1x, d2x, d3x, ay, d1y, d2y, d3y;
    act_eval_tc<AK>(act, z.v.x, ax, d1x, d2x, d3x);
    act_eval_tc<AK>(act, z.v.y, ay, d1y, d2y, d3y);
    a = mk2(ax, ay); d1 = mk2(d1x, d1y); d2 = mk2(d2x, d2y); d3 = mk2(d3x, d3y);
  }
}

// post-activation channels from pre-activation channels (z[0] value, z[1..N1], z[1+N1..]); T = float or P2
template <int N1, int N2, bool PURE, int AK, typename T>
__device__ __forceinline__ void chain_fwd(int act, const Chan<N1, N2>& ch, const T* z, T* h) {
  constexpr int M1 = (N1 > 0) ? N1 : 1;
  T a, d1, d2, d3;
  act_any<AK>(act, z[0], a, d1, d2, d3);
  h[0] = a;
#pragma unroll
  for (int i = 0; i < N1; ++i) h[1 + i] = d1 * z[1 + i];
#pragma unroll
  for (int s = 0; s < N2; ++s) {
    const T za = PURE ? z[1 + (s < N1 ? s : 0)] : pickT<T, M1>(z + 1, ch.sa[s]);
    const T zb = PURE ? za : pickT<T, M1>(z + 1, ch.sb[s]);
    h[1 + N1 + s] = vfma(d1, z[1 + N1 + s], d2 * za * zb);
  }
}

// adjoints of pre-activations from adjoints of post-activations; T = float or P2
template <int N1, int N2, bool PURE, int AK, typename T>
__device__ __forceinline__ void chain_bwd(int act, const Chan<N1, N2>& ch, const T* z, const T* hb, T* zb) {
  constexpr int M1 = (N1 > 0) ? N1 : 1;
  T a, d1, d2, d3;
  act_any<AK>(act, z[0], a, d1, d2, d3);
  T acc0 = d1 * hb[0];
#pragma unroll
  for (int i = 0; i < N1; ++i) {
    acc0 = vfma(d2 * z[1 + i], hb[1 + i], acc0);
    zb[1 + i] = d1 * hb[1 + i];
  }
#pragma unroll
  for (int s = 0; s < N2; ++s) {
    const T g = hb[1 + N1 + s];
    if (PURE) {
      const int i = s < N1 ? s : 0;
      const T za = z[1 + i];
      acc0 = vfma(vfma(d2, z[1 + N1 + s], d3 * za * za), g, acc0);
      zb[1 + i] = vfma((d2 * za) * vsplat<T>(2.f), g, zb[1 + i]);
    } else {
      const T za = pickT<T, M1>(z + 1, ch.sa[s]), zbb = pickT<T, M1>(z + 1, ch.sb[s]);
      acc0 = vfma(vfma(d2, z[1 + N1 + s], d3 * za * zbb), g, acc0);
      const T ga = d2 * zbb * g, gb2 = d2 * za * g;
#pragma unroll
      for (int i = 0; i < M1; ++i) {
        if (ch.sa[s] == i) zb[1 + i] = zb[1 + i] + ga;
        if (ch.sb[s] == i) zb[1 + i] = zb[1 + i] + gb2;
      }
    }
    zb[1 + N1 + s] = d1 * g;
  }
  zb[0] = acc0;
}

// sum over the 32 lanes of 4 per-lane values with 6 shuffles.  Every lane receives the total of
// element e = ((lane>>4)&1)*2 + ((lane>>3)&1); lanes with (lane & 7) == 0 act on it.
__device__ __forceinline__ float warp_reduce4(const float (&v)[4], int lane) {
#ifdef PINN_EXP_NO_RED
  return v[0] + v[1] + v[2] + v[3];
#endif
  const bool up16 = (lane & 16) != 0;
  float a0 = (up16 ? v[2] : v[0]) + __shfl_xor_sync(0xffffffffu, up16 ? v[0] : v[2], 16);
  float a1 = (up16 ? v[3] : v[1]) + __shfl_xor_sync(0xffffffffu, up16 ? v[1] : v[3], 16);
  const bool up8 = (lane & 8) != 0;
  float r = (up8 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, up8 ? a0 : a1, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
__device__ __forceinline__ int reduce4_elem(int lane) { return ((lane >> 4) & 1) * 2 + ((lane >> 3) & 1); }
// 2 values: 5 shuffles; every lane gets the total of element e = (lane >> 4) & 1; lanes with (lane & 15) == 0 act
__device__ __forceinline__ float warp_reduce2(const float (&v)[2], int lane) {
  const bool up16 = (lane & 16) != 0;
  float r = (up16 ? v[1] : v[0]) + __shfl_xor_sync(0xffffffffu, up16 ? v[0] : v[1], 16);
  r += __shfl_xor_sync(0xffffffffu, r, 8);
  r += __shfl_xor_sync(0xffffffffu, r, 4);
  r += __shfl_xor_sync(0xffffffffu, r, 2);
  r += __shfl_xor_sync(0xffffffffu, r, 1);
  return r;
}
__device__ __forceinline__ float warp_reduceg(const float (&v)[4], int lane) { return warp_reduce4(v, lane); }
__device__ __forceinline__ float warp_reduceg(const float (&v)[2], int lane) { return warp_reduce2(v, lane); }
__device__ __forceinline__ int reduceg_elem(int lane) { return GW == 4 ? reduce4_elem(lane) : ((lane >> 4) & 1); }
__device__ __forceinline__ bool reduceg_lead(int lane) { return GW == 4 ? ((lane & 7) == 0) : ((lane & 15) == 0); }

// phase timestamps for the timeline tool (scripts/tc_timeline.py): id in the high bits, clock in the low
// (only in -DPINN_DEBUG builds: libpinn_b200_debug.so; the product library carries no instrumentation)
template <typename CS>
__device__ __forceinline__ void dbg_mark(CS* cs, int id) {
#ifdef PINN_DEBUG
  if (cs->dbg && threadIdx.x == 0 && cs->dbg_n < 1000) {
    cs->dbg[cs->dbg_n++] = ((long long)id << 48) | (clock64() & 0xffffffffffffLL);
  }
#endif
}

struct Misc {   // carve-up of the misc region
  float *Xs, *taps, *tapbar, *scratch, *qws;
  double* tsum;
  uint64_t *bar_mma, *bar_ld;
  uint32_t* tmem_slot;
};
// mx_dim / mx_taps: largest point dimension / tap count over the problem's terms (the arrays are sized to them)
__device__ __forceinline__ Misc misc_of(uint8_t* m, int mx_dim, int mx_taps) {
  Misc r;
  r.Xs = reinterpret_cast<float*>(m);             m += mx_dim * kTcPts * 4;
  r.taps = reinterpret_cast<float*>(m);           m += mx_taps * kTcPts * 4;
  r.tapbar = reinterpret_cast<float*>(m);         m += mx_taps * kTcPts * 4;
  r.scratch = reinterpret_cast<float*>(m);        m += kTcMaxC * kTcPts * 4;
  r.qws = reinterpret_cast<float*>(m);            m += kTcPts * 4;
  r.tsum = reinterpret_cast<double*>(m);          m += PINN_MAX_TERMS * 8;
  r.bar_mma = reinterpret_cast<uint64_t*>(m);     m += 8;
  r.bar_ld = reinterpret_cast<uint64_t*>(m);      m += 8;
  r.tmem_slot = reinterpret_cast<uint32_t*>(m);
  return r;
}

// descriptor fields a network pass needs, read once from global memory into registers
template <int N1, int N2>
struct PassInfo {
  int L, TL, d_in, n1w, nL;
  int dir1[N1 > 0 ? N1 : 1];
  Chan<N1, N2> ch;
};
template <int N1, int N2>
__device__ __forceinline__ void load_pass(PassInfo<N1, N2>& pi, const DevNet& net, const DevChan& dc) {
  pi.L = net.n_layers; pi.TL = pi.L - 2; pi.d_in = net.dims[0]; pi.n1w = net.dims[1]; pi.nL = net.dims[pi.L - 1];
#pragma unroll
  for (int j = 0; j < N1; ++j) pi.dir1[j] = dc.dir1[j];
#pragma unroll
  for (int s = 0; s < N2; ++s) { pi.ch.sa[s] = dc.s_a[s]; pi.ch.sb[s] = dc.s_b[s]; }
}

__device__ __forceinline__ void wait_bar(uint64_t* bar, uint32_t& phase) {
  tc::mbar_wait(bar, phase);
  phase ^= 1u;
}

// issue a chain of nk MMAs D (+)= A_k * B_k; descriptors advance by a_step / b_step bytes per k-step
__device__ __forceinline__ void mma_chain(uint32_t d, uint64_t adesc, uint64_t bdesc, uint32_t a_step, uint32_t b_step, int nk,
                                          uint32_t idesc, uint32_t acc_first) {
  const uint64_t da = a_step >> 4, db = b_step >> 4;
#pragma unroll 1
  for (int k = 0; k < nk; ++k) {
    tc::mma_bf16(d, adesc, bdesc, idesc, (k > 0) ? 1u : acc_first);
    adesc += da;
    bdesc += db;
  }
}

// thread identity inside the CTA
struct Tid {
  int tid, warp, lane, q, hh, p;
  uint32_t lane_addr;
};
__device__ __forceinline__ Tid tid_of() {
  Tid t;
  t.tid = threadIdx.x; t.warp = t.tid >> 5; t.lane = t.tid & 31; t.q = t.warp & 3; t.hh = t.warp >> 2;
  t.p = t.q * 32 + t.lane;
  t.lane_addr = (uint32_t)(t.q * 32) << 16;
  return t;
}

#define PINN_TC_CASE(a1, a2, pu, CALL)                                                    \
  {                                                                                       \
    constexpr int A1 = a1, A2 = a2;                                                       \
    constexpr bool PU = pu;                                                               \
    if (_ak == 1) { constexpr int AK = 1; CALL; } else { constexpr int AK = 0; CALL; }   \
  }                                                                                       \
  break
// ak = 1 when every hidden layer of the network is tanh (fast branch-free activation), else 0
#define PINN_TC_DISPATCH(n1, n2, pure, ak, CALL)                              \
  do {                                                                        \
    const int _ak = (ak);                                                     \
    const int _key = ((n1) * 8 + (n2)) * 2 + ((pure) ? 1 : 0);                \
    switch (_key) {                                                           \
      case (0 * 8 + 0) * 2: case (0 * 8 + 0) * 2 + 1: PINN_TC_CASE(0, 0, true, CALL);   \
      case (1 * 8 + 0) * 2: case (1 * 8 + 0) * 2 + 1: PINN_TC_CASE(1, 0, true, CALL);   \
      case (2 * 8 + 0) * 2: case (2 * 8 + 0) * 2 + 1: PINN_TC_CASE(2, 0, true, CALL);   \
      case (3 * 8 + 0) * 2: case (3 * 8 + 0) * 2 + 1: PINN_TC_CASE(3, 0, true, CALL);   \
      case (4 * 8 + 0) * 2: case (4 * 8 + 0) * 2 + 1: PINN_TC_CASE(4, 0, true, CALL);   \
      case (1 * 8 + 1) * 2: case (1 * 8 + 1) * 2 + 1: PINN_TC_CASE(1, 1, true, CALL);   \
      case (2 * 8 + 1) * 2 + 1: PINN_TC_CASE(2, 1, true, CALL);               \
      case (2 * 8 + 1) * 2: PINN_TC_CASE(2, 1, false, CALL);                  \
      case (3 * 8 + 1) * 2 + 1: PINN_TC_CASE(3, 1, true, CALL);               \
      case (3 * 8 + 1) * 2: PINN_TC_CASE(3, 1, false, CALL);                  \
      case (2 * 8 + 2) * 2 + 1: PINN_TC_CASE(2, 2, true, CALL);               \
      case (2 * 8 + 2) * 2: PINN_TC_CASE(2, 2, false, CALL);                  \
      default: break;                                                         \
    }                                                                         \
  } while (0)

}  // namespace pinn
