// ffma_launch.cu -- gradient/loss reduction kernels and the host-callable launchers of the
// FFMA path.
#include <cuda_runtime.h>
#include "dev_types.h"

namespace pinn {

// ------------------------------------------------------------------------------------------
// Fixed-order reduction of the per-CTA partials; also turns the per-term sums into losses.
//   out_grad[i]      = sum_b partial[b][i]
//   out_terms[k]     = scale_k * sum_b term_sums[b][k]          (unweighted L_k)
//   out_total        = sum_k w_k * out_terms[k]
// When `packed` is non-null (multi-GPU), grad and the term losses are written contiguously
// into packed[0..n_theta+n_terms) for a single allreduce and finish_kernel unpacks.
template <typename real>
__global__ void reduce_kernel(const real* __restrict__ partial, long long stride, const double* __restrict__ term_sums, int nb,
                              long long n_theta, int n_terms, const ScaleW sw,
                              real* out_grad, real* out_terms, real* out_total, int want_grad) {
  // block = 32 gradient entries x 8 slices of the CTA partials; fixed summation order (slice-major)
  __shared__ real red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + tx;
  if (want_grad) {
    real s = real(0);
    if (i < n_theta) {
      const int chunk = (nb + 7) / 8;
      const int b0 = ty * chunk, b1 = (b0 + chunk < nb) ? b0 + chunk : nb;
      for (int b = b0; b < b1; ++b) s += partial[(long long)b * stride + i];
    }
    red[ty][tx] = s;
    __syncthreads();
    if (ty == 0 && i < n_theta) {
      real t = red[0][tx];
#pragma unroll
      for (int k = 1; k < 8; ++k) t += red[k][tx];
      out_grad[i] = t;
    }
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double tot = 0.0;
    for (int k = 0; k < n_terms; ++k) {
      double s = 0.0;
      for (int b = threadIdx.x; b < nb; b += 32) s += term_sums[(long long)b * PINN_MAX_TERMS + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
      double Lk = s * sw.scale[k];
      tot += Lk * sw.w[k];
      if (threadIdx.x == 0) out_terms[k] = real(Lk);
    }
    if (threadIdx.x == 0 && out_total) *out_total = real(tot);
  }
}

// gradient reduction fused with the Adam update: theta, m, v updated in place (single-GPU path)
template <typename real>
__global__ void reduce_adam_kernel(const real* __restrict__ partial, long long stride, const double* __restrict__ term_sums, int nb,
                                   long long n_theta, int n_terms, const ScaleW sw, real* theta, real* m, real* v,
                                   double lr_t, double beta1, double beta2, double eps_t, real* out_terms, real* out_total) {
  __shared__ real red[8][33];
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const long long i = (long long)blockIdx.x * 32 + tx;
  real s = real(0);
  if (i < n_theta) {
    const int chunk = (nb + 7) / 8;
    const int b0 = ty * chunk, b1 = (b0 + chunk < nb) ? b0 + chunk : nb;
    for (int b = b0; b < b1; ++b) s += partial[(long long)b * stride + i];
  }
  red[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && i < n_theta) {
    real g = red[0][tx];
#pragma unroll
    for (int k = 1; k < 8; ++k) g += red[k][tx];
    const double gd = (double)g;
    const double mi = beta1 * (double)m[i] + (1.0 - beta1) * gd;
    const double vi = beta2 * (double)v[i] + (1.0 - beta2) * gd * gd;
    m[i] = real(mi); v[i] = real(vi);
    // lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t), eps_t = eps * sqrt(1 - beta2^t)
    theta[i] = real((double)theta[i] - lr_t * mi / (sqrt(vi) + eps_t));
  }
  if (blockIdx.x == 0 && threadIdx.x < 32) {
    double tot = 0.0;
    for (int k = 0; k < n_terms; ++k) {
      double t = 0.0;
      for (int b = threadIdx.x; b < nb; b += 32) t += term_sums[(long long)b * PINN_MAX_TERMS + k];
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
      double Lk = t * sw.scale[k];
      tot += Lk * sw.w[k];
      if (threadIdx.x == 0) out_terms[k] = real(Lk);
    }
    if (threadIdx.x == 0) *out_total = real(tot);
  }
}

// after the allreduce of packed = [grad | term losses]: copy the gradient out and form total = sum_k w_k L_k
template <typename real>
__global__ void finish_kernel(const real* __restrict__ packed, long long n_grad, int n_terms, const ScaleW sw, real* out_grad,
                              real* out_terms, real* out_total) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (out_grad && i < n_grad) out_grad[i] = packed[i];
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    const real* packed_terms = packed + n_grad;
    double tot = 0.0;
    for (int k = 0; k < n_terms; ++k) {
      double Lk = (double)packed_terms[k];
      out_terms[k] = real(Lk);
      tot += Lk * sw.w[k];
    }
    *out_total = real(tot);
  }
}

// ---- host-callable launchers -------------------------------------------------------------------
size_t ffma_smem_bytes(int dtype, long long buf_elems, int w_area, bool bufs_smem) {
  size_t es = dtype == PINN_F64 ? 8 : 4;
  size_t n = (bufs_smem ? 2 * (size_t)buf_elems : 0) + (size_t)w_area + PINN_MAX_DIM * kTilePts +
             2 * PINN_MAX_TAPS * kTilePts + 2 * kTilePts;
  size_t bytes = n * es;
  bytes = (bytes + 7) & ~size_t(7);
  bytes += PINN_MAX_TERMS * sizeof(double);
  return bytes;
}

cudaError_t ffma_launch_float_smem(const FfmaArgs& a, int grid, size_t smem, cudaStream_t st);
cudaError_t ffma_launch_float_gmem(const FfmaArgs& a, int grid, size_t smem, cudaStream_t st);
cudaError_t ffma_launch_double_smem(const FfmaArgs& a, int grid, size_t smem, cudaStream_t st);
cudaError_t ffma_launch_double_gmem(const FfmaArgs& a, int grid, size_t smem, cudaStream_t st);

cudaError_t ffma_launch(int dtype, bool bufs_smem, const FfmaArgs& a, int grid, size_t smem, cudaStream_t st) {
  if (dtype == PINN_F64)
    return bufs_smem ? ffma_launch_double_smem(a, grid, smem, st) : ffma_launch_double_gmem(a, grid, smem, st);
  return bufs_smem ? ffma_launch_float_smem(a, grid, smem, st) : ffma_launch_float_gmem(a, grid, smem, st);
}

cudaError_t reduce_launch(int dtype, const void* partial, long long stride, const double* term_sums, int nb, long long n_theta,
                          int n_terms, const ScaleW& scale_w, void* out_grad, void* out_terms, void* out_total,
                          int want_grad, cudaStream_t st) {
  long long n = want_grad ? n_theta : 1;
  int blocks = (int)((n + 31) / 32);
  if (blocks < 1) blocks = 1;
  if (dtype == PINN_F64)
    reduce_kernel<double><<<blocks, 256, 0, st>>>((const double*)partial, stride, term_sums, nb, n_theta, n_terms, scale_w,
                                                   (double*)out_grad, (double*)out_terms, (double*)out_total,
                                                   want_grad);
  else
    reduce_kernel<float><<<blocks, 256, 0, st>>>((const float*)partial, stride, term_sums, nb, n_theta, n_terms, scale_w,
                                                  (float*)out_grad, (float*)out_terms, (float*)out_total, want_grad);
  return cudaGetLastError();
}

cudaError_t reduce_adam_launch(int dtype, const void* partial, long long stride, const double* term_sums, int nb, long long n_theta,
                               int n_terms, const ScaleW& sw, void* theta, void* m, void* v, double lr_t, double beta1,
                               double beta2, double eps_t, void* out_terms, void* out_total, cudaStream_t st) {
  int blocks = (int)((n_theta + 31) / 32);
  if (dtype == PINN_F64)
    reduce_adam_kernel<double><<<blocks, 256, 0, st>>>((const double*)partial, stride, term_sums, nb, n_theta, n_terms, sw,
                                                        (double*)theta, (double*)m, (double*)v, lr_t, beta1, beta2, eps_t,
                                                        (double*)out_terms, (double*)out_total);
  else
    reduce_adam_kernel<float><<<blocks, 256, 0, st>>>((const float*)partial, stride, term_sums, nb, n_theta, n_terms, sw,
                                                       (float*)theta, (float*)m, (float*)v, lr_t, beta1, beta2, eps_t,
                                                       (float*)out_terms, (float*)out_total);
  return cudaGetLastError();
}

cudaError_t finish_launch(int dtype, const void* packed, long long n_grad, int n_terms, const ScaleW& scale_w, void* out_grad,
                          void* out_terms, void* out_total, cudaStream_t st) {
  int blocks = (int)((n_grad + 255) / 256);
  if (blocks < 1) blocks = 1;
  if (dtype == PINN_F64)
    finish_kernel<double><<<blocks, 256, 0, st>>>((const double*)packed, n_grad, n_terms, scale_w, (double*)out_grad,
                                                   (double*)out_terms, (double*)out_total);
  else
    finish_kernel<float><<<blocks, 256, 0, st>>>((const float*)packed, n_grad, n_terms, scale_w, (float*)out_grad,
                                                  (float*)out_terms, (float*)out_total);
  return cudaGetLastError();
}

// max |g_i| and sum |g_i| of a gradient vector (one block; the vectors are at most a few MB)
template <typename real>
__global__ void __launch_bounds__(1024) grad_stats_kernel(const real* g, long long n, double* out) {
  __shared__ double smax[32], ssum[32];
  double mx = 0.0, sm = 0.0;
  for (long long i = threadIdx.x; i < n; i += 1024) {
    const double a = fabs((double)g[i]);
    mx = a > mx ? a : mx;
    sm += a;
  }
  for (int o = 16; o > 0; o >>= 1) {
    const double m2 = __shfl_xor_sync(0xffffffffu, mx, o);
    mx = m2 > mx ? m2 : mx;
    sm += __shfl_xor_sync(0xffffffffu, sm, o);
  }
  if ((threadIdx.x & 31) == 0) { smax[threadIdx.x >> 5] = mx; ssum[threadIdx.x >> 5] = sm; }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0, t = 0.0;
    for (int w = 0; w < 32; ++w) { m = smax[w] > m ? smax[w] : m; t += ssum[w]; }     // fixed order: reproducible
    out[0] = m;
    out[1] = n > 0 ? t / (double)n : 0.0;
  }
}

// ---- device-side StochasticTraining sampler ---------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter = (point index, group of 4 rows, draw), key = seed.  One thread per point
// writes its `dim` coordinates lb_r + (ub_r - lb_r) * u, u uniform in [0, 1) from the high 24 (float) / 53 (double) bits.
__device__ __forceinline__ void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}

struct SampleBox { double lb[PINN_MAX_DIM], ub[PINN_MAX_DIM]; };

template <typename real>
__global__ void __launch_bounds__(256) sample_uniform_kernel(real* pts, long long n, int dim, SampleBox box,
                                                              unsigned long long seed, unsigned long long draw,
                                                              const unsigned long long* draw_dev) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  if (draw_dev) draw += *draw_dev;       // device-side counter advanced by the fused kernel's tail (graph replays)
  for (int r0 = 0; r0 < dim; r0 += (sizeof(real) == 8 ? 2 : 4)) {
    uint32_t c[4] = {(uint32_t)p, (uint32_t)(p >> 32), (uint32_t)r0 ^ ((uint32_t)draw << 8), (uint32_t)(draw >> 24)};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    if (sizeof(real) == 8) {
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int r = r0 + j;
        if (r < dim) {
          const unsigned long long bits = ((unsigned long long)c[2 * j] << 32) | c[2 * j + 1];
          const double u = (double)(bits >> 11) * (1.0 / 9007199254740992.0);
          pts[p * dim + r] = (real)(box.lb[r] + (box.ub[r] - box.lb[r]) * u);
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int r = r0 + j;
        if (r < dim) {
          const float u = (float)(c[j] >> 8) * (1.0f / 16777216.0f);
          pts[p * dim + r] = (real)(box.lb[r] + (box.ub[r] - box.lb[r]) * (double)u);
        }
      }
    }
  }
}

// ---- device-side Latin hypercube sampler (QuasiRandomTraining's default sampling_alg = LatinHypercubeSample()) -----------
// Row r of point i is lb_r + (ub_r - lb_r) * (pi_r(i) + u) / n with pi_r a keyed permutation of [0, n) and u uniform in
// [0, 1): each of the n strata of every row holds exactly one point per draw.  pi_r is a 4-round Feistel network over
// ceil(log2 n) bits (each round XORs one half with a hash of the other: a bijection) with cycle walking back into [0, n):
// stateless, O(1) per point, no sort and no shuffle buffer.
__device__ __forceinline__ uint32_t mix32(uint32_t x) {
  x ^= x >> 16; x *= 0x85EBCA6Bu; x ^= x >> 13; x *= 0xC2B2AE35u; x ^= x >> 16;
  return x;
}
__device__ __forceinline__ uint32_t feistel_perm(uint32_t i, uint32_t n, uint32_t bits, uint32_t k0, uint32_t k1) {
  const uint32_t hb = bits >> 1, ob = bits - hb;
  const uint32_t ma = (1u << hb) - 1u, mb = (1u << ob) - 1u;      // bits >= 2 here
  uint32_t x = i;
  do {
    uint32_t a = x & ma, b = x >> hb;
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      a ^= mix32(b + k0 + 0x9E3779B9u * (uint32_t)(2 * r + 1)) & ma;
      b ^= mix32(a + k1 + 0x9E3779B9u * (uint32_t)(2 * r + 2)) & mb;
    }
    x = (b << hb) | a;
  } while (x >= n);
  return x;
}

template <typename real>
__global__ void __launch_bounds__(256) sample_lhs_kernel(real* pts, long long n, int dim, SampleBox box, unsigned long long seed,
                                                          unsigned long long draw, const unsigned long long* draw_dev) {
  const long long p = (long long)blockIdx.x * 256 + threadIdx.x;
  if (p >= n) return;
  if (draw_dev) draw += *draw_dev;
  uint32_t bits = 2;
  while ((1ull << bits) < (unsigned long long)n) ++bits;
  for (int r = 0; r < dim; ++r) {
    // per (row, draw) permutation key and per (point, row, draw) jitter from Philox
    uint32_t kc[4] = {(uint32_t)r, (uint32_t)draw, (uint32_t)(draw >> 32), 0x4C48535Fu};
    philox4x32_10(kc, (uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t c[4] = {(uint32_t)p, (uint32_t)(p >> 32), (uint32_t)r ^ ((uint32_t)draw << 8), (uint32_t)(draw >> 24) ^ 0x80000000u};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    const double u = (double)((((unsigned long long)c[0] << 32) | c[1]) >> 11) * (1.0 / 9007199254740992.0);
    double x;
    if (n == 1) {
      x = u;
    } else {
      const uint32_t j = feistel_perm((uint32_t)p, (uint32_t)n, bits, kc[0], kc[1]);
      x = ((double)j + u) / (double)n;
    }
    pts[p * dim + r] = (real)(box.lb[r] + (box.ub[r] - box.lb[r]) * x);
  }
}

cudaError_t sample_lhs_launch(int dtype, void* pts, long long n, int dim, const double* lb, const double* ub,
                              unsigned long long seed, unsigned long long draw, const unsigned long long* draw_dev,
                              cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  SampleBox box;
  for (int r = 0; r < PINN_MAX_DIM; ++r) { box.lb[r] = r < dim ? lb[r] : 0.0; box.ub[r] = r < dim ? ub[r] : 0.0; }
  const int blocks = (int)((n + 255) / 256);
  if (dtype == PINN_F64) sample_lhs_kernel<double><<<blocks, 256, 0, st>>>((double*)pts, n, dim, box, seed, draw, draw_dev);
  else sample_lhs_kernel<float><<<blocks, 256, 0, st>>>((float*)pts, n, dim, box, seed, draw, draw_dev);
  return cudaGetLastError();
}

cudaError_t sample_uniform_launch(int dtype, void* pts, long long n, int dim, const double* lb, const double* ub,
                                  unsigned long long seed, unsigned long long draw, const unsigned long long* draw_dev,
                                  cudaStream_t st) {
  if (n <= 0) return cudaSuccess;
  SampleBox box;
  for (int r = 0; r < PINN_MAX_DIM; ++r) { box.lb[r] = r < dim ? lb[r] : 0.0; box.ub[r] = r < dim ? ub[r] : 0.0; }
  const int blocks = (int)((n + 255) / 256);
  if (dtype == PINN_F64) sample_uniform_kernel<double><<<blocks, 256, 0, st>>>((double*)pts, n, dim, box, seed, draw, draw_dev);
  else sample_uniform_kernel<float><<<blocks, 256, 0, st>>>((float*)pts, n, dim, box, seed, draw, draw_dev);
  return cudaGetLastError();
}

cudaError_t grad_stats_launch(int dtype, const void* grad, long long n, double* out2, cudaStream_t st) {
  if (dtype == PINN_F64) grad_stats_kernel<double><<<1, 1024, 0, st>>>((const double*)grad, n, out2);
  else grad_stats_kernel<float><<<1, 1024, 0, st>>>((const float*)grad, n, out2);
  return cudaGetLastError();
}

}  // namespace pinn
