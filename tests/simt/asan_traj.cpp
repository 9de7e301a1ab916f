// TEST INFRASTRUCTURE: the trajectory kernels of soft-grip_b200/csrc/sg_traj.cuh under the SIMT emulator, built with
// -fsanitize=address,undefined and run over ragged shapes and grid sizes on exact-size heap buffers, so that an
// out-of-bounds access or undefined behaviour in the kernel source is caught on the CPU (tests/test_traj.py builds and
// runs it).  The launch geometry mirrors the host code of sg_api.cu.
#define SG_SIMT_EMU 1
#include <vector>
#include <cstdio>
#include "sg_traj.cuh"
using namespace sg;
template <typename T> void run(int W, int Tn, int nchan, int grid) {
  const long long rows = (long long)W * Tn, nel = rows * nchan;
  T* in = (T*)malloc(sizeof(T) * nel); T* out = (T*)malloc(sizeof(T) * nel);
  for (long long i = 0; i < nel; i++) in[i] = (T)(i % 97) * (T)0.25;
  double* mean = (double*)malloc(8 * nchan); double* sd = (double*)malloc(8 * nchan);
  for (int c = 0; c < nchan; c++) { mean[c] = 1.0; sd[c] = 2.0; }
  TrajNoiseArgs<T> A; A.in = in; A.out = out; A.nelem = nel; A.nchan = nchan; A.nacc = nchan / 2; A.sigma_acc = 0.7f; A.sigma_gyro = 0.06f;
  A.k0 = 1; A.k1 = 2; A.first_quad = 12345; A.mean = mean; A.stdev = sd;
  auto kn = sg_traj_noise_kernel<T>;
  simt::launch(kn, grid, 256, (size_t)3 * nchan * sizeof(T), A);        // exact-size dynamic shared memory: sigma | mean | 1 / std
  A.mean = nullptr; A.stdev = nullptr;
  simt::launch(kn, grid, 256, (size_t)3 * nchan * sizeof(T), A);
  // stats
  const int qpr = nchan / 4; int l = 32; while (l % qpr) l += 32; int block = l; while (block + l <= 384) block += l;
  TrajStatsArgs<T> S; S.in = in; S.nrows = rows; S.nchan = nchan; S.nblocks = grid;
  S.partial = (double*)malloc(sizeof(double) * grid * 2 * nchan); S.mean = mean; S.stdev = sd;
  auto k1 = sg_traj_stats_partial_kernel<T>; auto k2 = sg_traj_stats_final_kernel<T>;
  simt::launch(k1, grid, block, (size_t)block * 64, S);
  simt::launch(k2, 1, TRAJ_STATS_FINAL_THREADS, (size_t)(TRAJ_STATS_FINAL_THREADS / (2 * nchan)) * 2 * nchan * sizeof(double), S);
  // mask
  int* touch = (int*)malloc(sizeof(int) * rows); for (long long i = 0; i < rows; i++) touch[i] = (int)((i * 7) % 4) | ((i % 3) ? (1 << 30) : 0);
  int* fl = (int*)malloc(sizeof(int) * W); for (int w = 0; w < W; w++) fl[w] = 3;
  for (int mode = 0; mode < 2; mode++) {
    TrajMaskArgs<T> M; M.traj = out; M.touch = touch; M.nworlds = W; M.T_ = Tn; M.nchan = nchan; M.allf = 3; M.anybit = 1 << 30; M.mode = mode; M.fleft = mode ? fl : nullptr;
    auto km = sg_traj_mask_kernel<T>;
    simt::launch(km, grid, 256, 0, M);
  }
  free(in); free(out); free(mean); free(sd); free(S.partial); free(touch); free(fl);
}
int main() {
  const int shapes[][3] = {{1, 1, 4}, {3, 33, 12}, {21, 200, 12}, {5, 7, 64}, {40, 31, 24}, {2, 65, 36}, {9, 1, 52}};
  for (auto& s : shapes) for (int grid : {1, 3, 16}) { run<float>(s[0], s[1], s[2], grid); run<double>(s[0], s[1], s[2], grid); }
  printf("asan driver done\n");
  return 0;
}
