// ffma_inst.cu -- one explicit instantiation of the fused kernel per translation unit
// (compiled four times: {float,double} x {activation buffers in smem, in global}) so the
// build parallelises.  -DPINN_INST_REAL=float|double -DPINN_INST_BUFS=0|1
#include "ffma_kernel.cuh"

namespace pinn {

#define PINN_CAT2(a, b) a##b
#define PINN_CAT(a, b) PINN_CAT2(a, b)
#if PINN_INST_BUFS
#define PINN_BUFS_NAME smem
#else
#define PINN_BUFS_NAME gmem
#endif
#define PINN_LAUNCH_NAME PINN_CAT(PINN_CAT(PINN_CAT(ffma_launch_, PINN_INST_REAL), _), PINN_BUFS_NAME)

cudaError_t PINN_LAUNCH_NAME(const FfmaArgs& a, int grid, size_t smem, cudaStream_t st) {
  auto k = ffma_loss_grad_kernel<PINN_INST_REAL, (PINN_INST_BUFS != 0)>;
  static size_t granted[64] = {0};
  cudaError_t e = ensure_dynamic_smem(k, smem, granted);
  if (e != cudaSuccess) return e;
  return launch_fused_kernel(k, a, grid, kThreads, smem, st, a.tail.state != nullptr);
}

}  // namespace pinn
