// altro_b200_large_mma.cu — translation unit of the tensor-instruction large-state kernel
// (large_mma.cuh).  Regular flags (FMA contraction on): this kernel agrees with the oracle to
// rounding; the exact-order kernel is in altro_b200_large.cu.
#include "large_mma.cuh"

using namespace altro_b200;

cudaError_t altro_b200_launch_solve_large_mma_32_8(const SolverParams& P, int mode, int sm_count, cudaStream_t st) {
  const int smem = MmaLayout::total * static_cast<int>(sizeof(double));
  cudaError_t e = cudaFuncSetAttribute(k_solve_large_mma<32, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return e;
  const int per_round = sm_count > 0 ? sm_count : 148;
  const int grid = P.B < per_round ? P.B : per_round;
  k_solve_large_mma<32, 8><<<grid, kMmaThreads, smem, st>>>(P, mode);
  return cudaGetLastError();
}
