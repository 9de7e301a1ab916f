// Dev probe (not part of the product): tensor memory as per-thread scratch.  Every thread of a 512-thread CTA owns
// 128 columns x 4 B of its TMEM lane (warp w: lane quarter w % 4, column block w / 4); the probe runs the access
// pattern of the equality sweep -- load a (u, n) pair, a few flops, store u, warp barrier -- against tensor memory and
// against shared memory, checks the results against each other and prints cycles per step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/tmem_probe scripts/tmem_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_ld2(uint32_t ta, float& x, float& y) {
  uint32_t a, b;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(ta));
  x = __uint_as_float(a); y = __uint_as_float(b);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_st1(uint32_t ta, float x) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" :: "r"(ta), "r"(__float_as_uint(x)));
}
__device__ __forceinline__ void tm_st2(uint32_t ta, float x, float y) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" :: "r"(ta), "r"(__float_as_uint(x)), "r"(__float_as_uint(y)));
}

constexpr int NSTEP = 48;

template <int MODE>   // 0: tensor memory, 1: shared memory
__global__ void __launch_bounds__(512, 1) probe(float* out, long long* cyc, int sweeps) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t tm_base;
  const int w = threadIdx.x >> 5;
  if (w == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" :: "r"((uint32_t)__cvta_generic_to_shared(&tm_base)));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t ta = tm_base + ((uint32_t)((w & 3) * 32) << 16) + (uint32_t)((w >> 2) * 128);
  float2* sm = reinterpret_cast<float2*>(smem) + (size_t)threadIdx.x * (NSTEP + 1);   // odd stride in 8-byte units
  const int gid = blockIdx.x * blockDim.x + threadIdx.x;
  for (int c = 0; c < NSTEP; c++) {
    const float u = 1e-3f * (float)((gid * 31 + c * 7) % 1000), n = -0.25f - 1e-4f * (float)(c + (gid & 15));
    if (MODE == 0) tm_st2(ta + 2 * c, u, n); else sm[c] = make_float2(u, n);
  }
  if (MODE == 0) tm_wait_st();
  __syncwarp();
  float a1 = 0.5f + 1e-3f * (float)(gid & 63), acc = 0;
  const long long t0 = clock64();
  for (int it = 0; it < sweeps; it++) {
    float u, n;
    if (MODE == 0) tm_ld2(ta, u, n); else { const float2 v = sm[0]; u = v.x; n = v.y; }
#pragma unroll 4
    for (int c = 0; c < NSTEP; c++) {
      if (MODE == 0) tm_wait_ld();
      const float uc = u, nc = n;
      if (c + 1 < NSTEP) { if (MODE == 0) tm_ld2(ta + 2 * (c + 1), u, n); else { const float2 v = sm[c + 1]; u = v.x; n = v.y; } }
      const float res = a1 + uc;
      const float dl = res * nc;
      acc += dl * res;
      a1 += 0.5f * dl;
      const float un = -a1 * 0.999f;
      if (MODE == 0) tm_st1(ta + 2 * c, un); else sm[c].x = un;
      __syncwarp();
    }
    if (MODE == 0) tm_wait_st();
  }
  const long long t1 = clock64();
  float s = acc + a1;
  for (int c = 0; c < NSTEP; c++) {
    float u, n;
    if (MODE == 0) { tm_ld2(ta + 2 * c, u, n); tm_wait_ld(); } else { u = sm[c].x; n = sm[c].y; }
    s += u * (float)(c + 1) + n;
  }
  out[gid] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (w == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" :: "r"(tm_base));
}

int main() {
  const int grid = 148, sweeps = 2000;
  for (int block : {32, 128, 512}) {
    float *o0, *o1; long long *c0, *c1;
    cudaMalloc(&o0, grid * 512 * 4); cudaMalloc(&o1, grid * 512 * 4); cudaMalloc(&c0, grid * 8); cudaMalloc(&c1, grid * 8);
    cudaMemset(o0, 0, grid * 512 * 4); cudaMemset(o1, 0, grid * 512 * 4);
    const size_t smem = (size_t)512 * (NSTEP + 1) * 8;
    cudaFuncSetAttribute(probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    probe<0><<<grid, block, smem>>>(o0, c0, sweeps);
    probe<1><<<grid, block, smem>>>(o1, c1, sweeps);
    cudaError_t e = cudaGetLastError(); if (e == cudaSuccess) e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
    static float h0[148 * 512], h1[148 * 512]; long long hc0[148], hc1[148];
    cudaMemcpy(h0, o0, sizeof h0, cudaMemcpyDeviceToHost); cudaMemcpy(h1, o1, sizeof h1, cudaMemcpyDeviceToHost);
    cudaMemcpy(hc0, c0, sizeof hc0, cudaMemcpyDeviceToHost); cudaMemcpy(hc1, c1, sizeof hc1, cudaMemcpyDeviceToHost);
    int bad = 0;
    for (int b = 0; b < grid; b++) for (int t = 0; t < block; t++) if (h0[b * block + t] != h1[b * block + t]) bad++;
    printf("block %3d: tensor memory %.1f cycles/step, shared memory %.1f cycles/step, mismatches %d of %d (sample %g %g)\n", block,
           (double)hc0[0] / (sweeps * NSTEP), (double)hc1[0] / (sweeps * NSTEP), bad, grid * block, h0[5], h1[5]);
  }
  return 0;
}
