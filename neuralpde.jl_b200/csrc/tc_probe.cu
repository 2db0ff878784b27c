// tc_probe.cu -- diagnostic: run one chain of tcgen05.mma instructions on caller-supplied
// shared-memory images and descriptor fields and return the fp32 accumulator.  Used by
// scripts/tc_probe.py and scripts/tc_probe_wide.py to pin the descriptor conventions (K-major / MN-major, LBO /
// SBO meaning, k-step advance) against numpy before the fused kernel relies on them.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>
#include "tc_prims.cuh"

namespace pinn {

struct ProbeArgs {
  const uint8_t* a_img; const uint8_t* b_img;
  uint32_t a_bytes, b_bytes;
  uint32_t a_off, b_off, a_lbo, a_sbo, b_lbo, b_sbo, a_step, b_step, n_ksteps, idesc, n_cols;
  float* out;   // [128][n_cols]
};

__global__ void __launch_bounds__(128, 1) tc_probe_kernel(const ProbeArgs pa) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint64_t bar;
  __shared__ uint32_t tmem_base_s;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((pa.a_bytes + 1023u) & ~1023u);
  for (uint32_t i = tid; i < pa.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(sa)[i] = reinterpret_cast<const uint4*>(pa.a_img)[i];
  for (uint32_t i = tid; i < pa.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(sb)[i] = reinterpret_cast<const uint4*>(pa.b_img)[i];
  if (tid == 0) { tc::mbar_init(&bar, 1); tc::fence_barrier_init(); }
  if (warp == 0) tc::tmem_alloc<256>(&tmem_base_s);
  tc::fence_async_smem();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (tid == 0) {
    for (uint32_t k = 0; k < pa.n_ksteps; ++k) {
      uint64_t da = tc::make_desc(tc::smem_u32(sa) + pa.a_off + k * pa.a_step, pa.a_lbo, pa.a_sbo);
      uint64_t db = tc::make_desc(tc::smem_u32(sb) + pa.b_off + k * pa.b_step, pa.b_lbo, pa.b_sbo);
      tc::mma_bf16(tmem, da, db, pa.idesc, k > 0 ? 1u : 0u);
    }
    tc::mma_commit(&bar);
  }
  tc::mbar_wait(&bar, 0);
  tc::tc_fence_after();
  for (uint32_t c0 = 0; c0 < pa.n_cols; c0 += 8) {
    float v[8];
    tc::tmem_ld8(tmem + ((uint32_t)(warp * 32) << 16) + c0, v);
    tc::tmem_ld_wait();
    for (int i = 0; i < 8; ++i) pa.out[(warp * 32 + lane) * pa.n_cols + c0 + i] = v[i];
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 0) tc::tmem_dealloc<256>(tmem);
}

}  // namespace pinn

extern "C" int pinn_debug_mma_probe(const void* a_img, uint32_t a_bytes, const void* b_img, uint32_t b_bytes,
                                    const uint32_t* p /* a_off,b_off,a_lbo,a_sbo,b_lbo,b_sbo,a_step,b_step,n_ksteps,idesc,n_cols */,
                                    float* host_out) {
  using namespace pinn;
  ProbeArgs pa;
  memset(&pa, 0, sizeof pa);
  uint8_t *da = nullptr, *db = nullptr;
  float* dout = nullptr;
  const uint32_t n_cols = p[10];
  if (cudaMalloc(&da, a_bytes) != cudaSuccess || cudaMalloc(&db, b_bytes) != cudaSuccess ||
      cudaMalloc(&dout, 128 * n_cols * sizeof(float)) != cudaSuccess) return 1;
  cudaMemcpy(da, a_img, a_bytes, cudaMemcpyHostToDevice);
  cudaMemcpy(db, b_img, b_bytes, cudaMemcpyHostToDevice);
  pa.a_img = da; pa.b_img = db; pa.a_bytes = a_bytes; pa.b_bytes = b_bytes;
  pa.a_off = p[0]; pa.b_off = p[1]; pa.a_lbo = p[2]; pa.a_sbo = p[3]; pa.b_lbo = p[4]; pa.b_sbo = p[5];
  pa.a_step = p[6]; pa.b_step = p[7]; pa.n_ksteps = p[8]; pa.idesc = p[9]; pa.n_cols = n_cols; pa.out = dout;
  size_t smem = ((a_bytes + 1023u) & ~1023u) + ((b_bytes + 1023u) & ~1023u) + 1024;
  cudaFuncSetAttribute(tc_probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  tc_probe_kernel<<<1, 128, smem>>>(pa);
  cudaError_t e = cudaDeviceSynchronize();
  int rc = 0;
  if (e != cudaSuccess) rc = 2;
  else cudaMemcpy(host_out, dout, 128 * n_cols * sizeof(float), cudaMemcpyDeviceToHost);
  cudaFree(da); cudaFree(db); cudaFree(dout);
  return rc;
}
