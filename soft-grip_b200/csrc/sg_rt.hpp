// sg_rt.hpp -- the one place that decides how device code is launched: nvcc + the CUDA runtime for the product,
// or (tests only, -DSG_SIMT_EMU) g++ + the SIMT emulator of tests/simt so the same kernel source runs on the CPU.
#pragma once
#ifdef SG_SIMT_EMU
#include "simt.h"
#include "cuda_shim.h"
#define SG_LAUNCH(kp, grid, block, smem, stream, arg) simt::launch(kp, (unsigned)(grid), (unsigned)(block), (size_t)(smem), arg)
#define SG_SHARED_BYTES(name) unsigned char* name = simt::smem_ptr()
#else
#include <cuda_runtime.h>
#define SG_LAUNCH(kp, grid, block, smem, stream, arg) kp<<<grid, block, smem, stream>>>(arg)
#define SG_SHARED_BYTES(name) extern __shared__ __align__(16) unsigned char name[]
#endif
