// altro_b200_large.cu — translation unit of the large-state path (large.cuh).
// Built with -fmad=false: every multiply and add rounds separately, in the reference's order,
// so this path reproduces the CPU oracle bit for bit (the C5 problem is ill-conditioned enough
// that the LLT success decisions depend on the last bit).
#include "large.cuh"

using namespace altro_b200;

cudaError_t altro_b200_launch_solve_large_32_8(const SolverParams& P, int mode, cudaStream_t st) {
  const int smem = sizeof(LargeSmem<32, 8>);
  cudaError_t e = cudaFuncSetAttribute(k_solve_large<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  k_solve_large<32, 8><<<P.B, kLargeThreads, smem, st>>>(P, mode);
  return cudaGetLastError();
}
