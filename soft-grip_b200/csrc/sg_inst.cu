// sg_inst.cu -- one instantiation of the step kernel: -DSG_INST_T=float|double -DSG_INST_LPW=4|8|16|32
#include "sg_rt.hpp"
#include "sg_launch.hpp"

namespace sg {

template <typename T, int LPW>
int k2_launch(const KArgs2<T>& K, int grid, int block, size_t smem, void* stream) {
  auto kp = sg_step_kernel2<T, LPW>;
  SG_LAUNCH(kp, grid, block, smem, (cudaStream_t)stream, K);
  return (int)cudaGetLastError();
}

template <typename T, int LPW>
int k2_configure(int block, size_t smem, int* per_sm) {
  auto kp = sg_step_kernel2<T, LPW>;
  cudaError_t e = cudaFuncSetAttribute(kp, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return (int)e;
  e = cudaFuncSetAttribute(kp, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
  if (e != cudaSuccess) return (int)e;
  return (int)cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, kp, block, smem);
}

template int k2_launch<SG_INST_T, SG_INST_LPW>(const KArgs2<SG_INST_T>&, int, int, size_t, void*);
template int k2_configure<SG_INST_T, SG_INST_LPW>(int, size_t, int*);

}  // namespace sg
