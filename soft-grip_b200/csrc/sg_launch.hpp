// sg_launch.hpp -- host launchers of sg_step_kernel2<T, LPW>; each (T, LPW) pair is instantiated in its own
// translation unit (sg_inst.cu compiled with -DSG_INST_T=.. -DSG_INST_LPW=..) so the build runs in parallel.
#pragma once
#include <stddef.h>

#include "sg_kernels2.cuh"

namespace sg {
// returns the CUDA error code of the launch (0 = ok)
template <typename T, int LPW> int k2_launch(const KArgs2<T>& K, int grid, int block, size_t smem, void* stream);
// sets the dynamic shared-memory limit / carve-out and reports resident CTAs per SM
template <typename T, int LPW> int k2_configure(int block, size_t smem, int* per_sm);
}  // namespace sg
