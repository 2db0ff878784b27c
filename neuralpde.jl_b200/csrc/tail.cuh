// tail.cuh -- the tail every fused loss+gradient kernel ends with: the gradient reduction, the optimizer step and the
// multi-GPU sum run INSIDE the fused kernel instead of in follow-up launches (reference: the Zygote gradient returned to
// Optimization.jl, src/discretize.jl:778, followed by Optimisers.Adam on the host; there is no multi-GPU counterpart).
//
//   1. grid barrier        all CTAs of the launch are co-resident (grid <= SM count, cooperative launch); a
//                          self-resetting generation barrier in global memory, so CUDA-graph replays need no host state
//   2. slice reduction     CTA b sums gradient entries [b*S, (b+1)*S) over the per-CTA partials in a FIXED order
//                          (bitwise reproducible for a given grid); CTA 0 turns the per-CTA term sums into losses
//   3. one-shot allreduce  (nranks > 1) the slice goes to this rank's symmetric buffer; per-slice flags are written
//                          straight into every peer's memory over NVLink (st.release.sys), each CTA waits for the
//                          same slice of every peer and adds the peers' slices in rank order (ld.relaxed.sys over
//                          NVLink) -- one hop, no extra launch, identical bits on every rank
//   4. consume             write the gradient, or apply Adam in place (theta, m, v resident on the device; bias
//                          correction from a device-side step counter, so a captured graph can be replayed)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include "dev_types.h"

namespace pinn {

// launch with the cooperative attribute when the kernel ends with the grid-wide tail (all CTAs must be co-resident)
template <typename K, typename A>
static cudaError_t launch_fused_kernel(K kernel, const A& a, int grid, int threads, size_t smem, cudaStream_t st, bool coop) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3((unsigned)threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  static const bool no_coop = [] { const char* v = getenv("PINN_B200_COOP"); return v && v[0] == '0'; }();   // measurement aid
  at[0].val.cooperative = (coop && !no_coop) ? 1 : 0;
  cfg.attrs = at; cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kernel, a);
}

__device__ __forceinline__ unsigned int ld_acquire_gpu(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ float ld_relaxed_sys(const float* p) {
  float v;
  asm volatile("ld.relaxed.sys.global.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double* p) {
  double v;
  asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ float4 ld_relaxed_sys_v(const float4* p) {
  float4 v;
  asm volatile("ld.relaxed.sys.global.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ double2 ld_relaxed_sys_v(const double2* p) {
  double2 v;
  asm volatile("ld.relaxed.sys.global.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long tail_now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// spin until pred() holds; a protocol failure (lost peer, CTAs not co-resident) traps after timeout_ns instead of hanging
template <typename Pred>
__device__ __forceinline__ void tail_spin(Pred pred, unsigned long long timeout_ns) {
  unsigned long long t0 = 0;
  for (unsigned int it = 0;; ++it) {
    if (pred()) return;
    if ((it & 1023u) == 1023u) {
      const unsigned long long t = tail_now_ns();
      if (t0 == 0) t0 = t;
      else if (t - t0 > timeout_ns) __trap();
    }
  }
}

template <typename real> struct Vec16;
template <> struct Vec16<float> { typedef float4 type; };
template <> struct Vec16<double> { typedef double2 type; };

// NT = threads per CTA.  red: NT * (16 / sizeof(real)) scalars of 16-byte aligned shared memory.  partial:
// [gridDim.x][stride] (stride a multiple of 4 scalars); term_sums: [gridDim.x][PINN_MAX_TERMS] (this CTA's row already
// written).  Must be called by every thread of every CTA of the launch.
template <typename real, int NT>
__device__ __noinline__ void fused_tail(const TailArgs& ta, const real* partial, long long stride, const double* term_sums,
                                        long long n_theta, int n_terms, int want_grad, real* red) {
  constexpr int V = 16 / (int)sizeof(real);   // scalars per 16-byte load: the reduction is bound by bytes in flight
  constexpr int EW = 32 * V;                  // gradient entries per pass (one warp of 16-byte vectors)
  constexpr int NG = NT / 32;                 // partial-row groups summed concurrently (one warp each)
  typedef typename Vec16<real>::type vec_t;
  __shared__ unsigned int s_step;
  __shared__ unsigned long long s_t;
  TailState* st = ta.state;
  const int tid = threadIdx.x, nb = gridDim.x, bid = blockIdx.x;

  // ---- 1. grid barrier ----------------------------------------------------------------------------------------------
  __syncthreads();
  if (tid == 0) {
    const unsigned int step = *reinterpret_cast<volatile unsigned int*>(&st->step);
    const unsigned long long t_adam = *reinterpret_cast<volatile unsigned long long*>(&st->adam_t) + 1ull;
    const unsigned int gen = *reinterpret_cast<volatile unsigned int*>(&st->gen);
    s_step = step;
    s_t = t_adam;
    __threadfence();
    if (atomicAdd(&st->count, 1u) == (unsigned int)(nb - 1)) {
      st->count = 0u;
      st->step = step + 1u;
      if (ta.adam_theta && want_grad) st->adam_t = t_adam;
      if (ta.bump_draw) st->draw = st->draw + 1ull;
      __threadfence();
      atomicAdd(&st->gen, 1u);
    } else {
      tail_spin([&] { return ld_acquire_gpu(&st->gen) != gen; }, ta.timeout_ns);
    }
    __threadfence();
  }
  __syncthreads();
  const unsigned int step1 = s_step + 1u;          // flag value of this step (0 = never signalled)
  const bool multi = ta.nranks > 1;
  real* mybuf = multi ? reinterpret_cast<real*>(ta.peer_buf[s_step & 1u][ta.rank]) : nullptr;

  // ---- 2. term losses (CTA 0, one warp per term): L_k = scale_k * sum_b term_sums[b][k], fixed order ---------------------
  if (bid == 0) {
    double* sL = reinterpret_cast<double*>(red);          // n_terms doubles (<= 32 * 8 bytes: fits NT scalars)
    const int warp = tid >> 5, lane = tid & 31;
    for (int k = warp; k < n_terms; k += NT / 32) {
      double v[(kTailSlots + 31) / 32];
#pragma unroll
      for (int j = 0; j < (kTailSlots + 31) / 32; ++j) {
        const int b = lane + 32 * j;
        v[j] = (b < nb) ? __ldcg(&term_sums[(long long)b * PINN_MAX_TERMS + k]) : 0.0;
      }
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < (kTailSlots + 31) / 32; ++j) s += v[j];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      if (lane == 0) sL[k] = s * ta.sw.scale[k];
    }
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int k = 0; k < n_terms; ++k) {
        const double Lk = sL[k];
        tot += Lk * ta.sw.w[k];
        if (multi) reinterpret_cast<double*>(reinterpret_cast<char*>(mybuf) + ta.terms_off)[k] = Lk;
        else reinterpret_cast<real*>(ta.out_terms)[k] = real(Lk);
      }
      if (!multi && ta.out_total) *reinterpret_cast<real*>(ta.out_total) = real(tot);
    }
    __syncthreads();
  }

  // Adam coefficients (device-side step counter: graph replays advance it)
  double lr_t = 0.0, eps_t = 0.0;
  const bool adam = ta.adam_theta != nullptr && want_grad;
  if (adam) {
    const double t = (double)s_t;
    const double c1 = 1.0 - pow(ta.adam_b1, t), c2 = sqrt(1.0 - pow(ta.adam_b2, t));
    lr_t = ta.adam_lr * c2 / c1;                    // lr * sqrt(1 - beta2^t) / (1 - beta1^t)
    eps_t = ta.adam_eps * c2;
  }
  auto consume = [&](long long i, real g) {
    if (adam) {
      real* th = reinterpret_cast<real*>(ta.adam_theta);
      real* m = reinterpret_cast<real*>(ta.adam_m);
      real* v = reinterpret_cast<real*>(ta.adam_v);
      const double gd = (double)g;
      const double mi = ta.adam_b1 * (double)m[i] + (1.0 - ta.adam_b1) * gd;
      const double vi = ta.adam_b2 * (double)v[i] + (1.0 - ta.adam_b2) * gd * gd;
      m[i] = real(mi); v[i] = real(vi);
      th[i] = real((double)th[i] - lr_t * mi / (sqrt(vi) + eps_t));
      if (ta.out_grad) reinterpret_cast<real*>(ta.out_grad)[i] = g;
    } else {
      reinterpret_cast<real*>(ta.out_grad)[i] = g;
    }
  };

  // ---- 3. slice reduction over the per-CTA partials ----------------------------------------------------------------------
  // warp g adds rows g, g + NG, ... of a 32-vector-wide window (16-byte L2 loads, all in flight at once), the NG row
  // groups are then combined through shared memory in group order: a fixed summation order for a given grid
  long long S = (n_theta + nb - 1) / nb;
  S = (S + V - 1) / V * V;
  const long long i0 = (long long)bid * S;
  const long long i1 = (i0 + S < n_theta) ? i0 + S : n_theta;
  const int lane = tid & 31, g = tid >> 5;
  if (want_grad) {
    for (long long base = i0; base < i1; base += EW) {
      const long long i = base + (long long)V * lane;
      real acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = real(0);
      if (i < i1) {
        const real* col = partial + i;
#pragma unroll 8
        for (int b = g; b < nb; b += NG) {
          const vec_t v = __ldcg(reinterpret_cast<const vec_t*>(col + (long long)b * stride));
          const real* pv = reinterpret_cast<const real*>(&v);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += pv[j];
        }
      }
      *reinterpret_cast<vec_t*>(red + (size_t)(g * 32 + lane) * V) = *reinterpret_cast<vec_t*>(acc);
      __syncthreads();
      if (tid < EW && base + tid < i1) {
        real t = red[tid];
#pragma unroll
        for (int k = 1; k < NG; ++k) t += red[k * EW + tid];
        if (multi) mybuf[base + tid] = t; else consume(base + tid, t);
      }
      __syncthreads();
    }
  }
  if (!multi) return;

  // ---- 4. one-shot allreduce over peer memory: signal slice `bid` to every peer, wait for theirs, add in rank order -------
  // (the release store orders this CTA's slice, made visible to the signalling thread by the barrier, before the flag:
  //  no CTA-wide system fence)
  __syncthreads();
  if (tid < ta.nranks && tid != ta.rank) {
    st_release_sys(ta.peer_flags[tid] + (size_t)ta.rank * kTailSlots + bid, step1);
    const unsigned int* mine = ta.peer_flags[ta.rank] + (size_t)tid * kTailSlots + bid;
    tail_spin([&] { return (int)(ld_acquire_sys(mine) - step1) >= 0; }, ta.timeout_ns);
  }
  __syncthreads();
  if (want_grad) {
    for (long long i = i0 + (long long)V * tid; i < i1; i += (long long)V * NT) {
      real acc[V];
#pragma unroll
      for (int j = 0; j < V; ++j) acc[j] = real(0);
#pragma unroll
      for (int r = 0; r < kMaxRanks; ++r)          // all peers' 16-byte loads in flight together, added in rank order
        if (r < ta.nranks) {
          const vec_t v = ld_relaxed_sys_v(reinterpret_cast<const vec_t*>(reinterpret_cast<const real*>(ta.peer_buf[s_step & 1u][r]) + i));
          const real* pv = reinterpret_cast<const real*>(&v);
#pragma unroll
          for (int j = 0; j < V; ++j) acc[j] += pv[j];
        }
#pragma unroll
      for (int j = 0; j < V; ++j)
        if (i + j < i1) consume(i + j, acc[j]);
    }
  }
  if (bid == 0 && tid == 0) {
    double tot = 0.0;
    for (int k = 0; k < n_terms; ++k) {
      double Lk = 0.0;
      for (int r = 0; r < ta.nranks; ++r)
        Lk += ld_relaxed_sys(reinterpret_cast<const double*>(reinterpret_cast<const char*>(ta.peer_buf[s_step & 1u][r]) + ta.terms_off) + k);
      reinterpret_cast<real*>(ta.out_terms)[k] = real(Lk);
      tot += Lk * ta.sw.w[k];
    }
    if (ta.out_total) *reinterpret_cast<real*>(ta.out_total) = real(tot);
  }
}

}  // namespace pinn
