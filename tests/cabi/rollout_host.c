/* A plain-C caller of libsoftgrip.so (include/softgrip.h): no Python, no torch, no CUDA headers.
 *
 *   rollout_host <model.sgm> <nworlds> <out.bin>
 *
 * Loads a compiled model blob, creates a batch of fp32 worlds, runs the squeeze episode of the reference driver
 * (ref: create_dataset.py:33-60: 40 settle rows, close_hand, 160 rows with toggle_grip at i = 80; sim_start 1, sim_step 7)
 * through sg_batch_rollout_host with HOST buffers only, and writes [nworlds][200][12] floats + [nworlds] status ints.
 * Exit codes: 0 ok, 3 = the library reported "no CUDA device" (what a CPU-only box must get: there is no CPU path),
 * 1 = any other failure.  Built and run by tests/test_abi.py. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "softgrip.h"

static int die(const char* what) {
  const char* e = sg_last_error();
  fprintf(stderr, "%s: %s\n", what, e ? e : "?");
  return (e && strstr(e, "no CUDA device")) ? 3 : 1;
}

int main(int argc, char** argv) {
  if (argc < 4) { fprintf(stderr, "usage: %s model.sgm nworlds out.bin\n", argv[0]); return 1; }
  FILE* f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 1; }
  fseek(f, 0, SEEK_END);
  long n = ftell(f);
  fseek(f, 0, SEEK_SET);
  void* blob = malloc((size_t)n);
  if (fread(blob, 1, (size_t)n, f) != (size_t)n) { fclose(f); return 1; }
  fclose(f);
  const int W = atoi(argv[2]);

  sg_model* m = NULL;
  if (sg_model_load(blob, (size_t)n, &m)) return die("sg_model_load");
  sg_info info;
  if (sg_model_info(m, &info)) return die("sg_model_info");
  printf("model: nv %d nshell %d neq %d nu %d nsensordata %d levels %d\n", info.nv, info.nshell, info.neq, info.nu, info.nsensordata, info.nlevels);

  /* ManEnv.joint_ids = range(11, 64), tendon_ids = [0] (ref: manenv.py:12-13) */
  int* mask = (int*)calloc((size_t)info.nv, sizeof(int));
  for (int j = 11; j < 64 && j < info.nv; j++) mask[j] = 1;
  if (sg_model_set_stiffness_targets(m, mask, 1)) return die("sg_model_set_stiffness_targets");

  sg_batch* b = NULL;
  if (sg_batch_create(m, W, 0, 32, &b)) { int rc = die("sg_batch_create"); sg_model_destroy(m); return rc; }

  enum { SETTLE = 40, ITER = 160, DIV = 80, T = SETTLE + ITER };
  int ev[T];
  double val[T * 2];
  memset(ev, 0, sizeof(ev));
  memset(val, 0, sizeof(val));
  int closing = 1;
  for (int i = 0; i < ITER; i++) {
    const int t = SETTLE + i;
    if (i == 0) { ev[t] = 1; val[2 * t] = val[2 * t + 1] = -0.2; }                     /* close_hand */
    else if (i % DIV == 0) { closing = !closing; ev[t] = 1; val[2 * t] = val[2 * t + 1] = closing ? -0.2 : 0.2; }  /* toggle_grip */
  }
  sg_schedule sc = {1, 7, T, ev, val};

  double* k = (double*)malloc(sizeof(double) * (size_t)W);
  for (int w = 0; w < W; w++) k[w] = W > 1 ? 300.0 + (1100.0 * w) / (W - 1) : 850.0;
  float* traj = (float*)malloc(sizeof(float) * (size_t)W * T * (size_t)info.nsensordata);
  int* status = (int*)malloc(sizeof(int) * (size_t)W);
  if (sg_batch_rollout_host(b, &sc, k, traj, NULL, status)) return die("sg_batch_rollout_host");

  FILE* o = fopen(argv[3], "wb");
  if (!o) { perror(argv[3]); return 1; }
  fwrite(traj, sizeof(float), (size_t)W * T * (size_t)info.nsensordata, o);
  fwrite(status, sizeof(int), (size_t)W, o);
  fclose(o);
  double chk = 0;
  for (size_t i = 0; i < (size_t)W * T * (size_t)info.nsensordata; i++) chk += traj[i] < 0 ? -traj[i] : traj[i];
  printf("rollout: %d worlds x %d rows, sum|x| = %.6e, launches %lld\n", W, T, chk, sg_batch_launch_count(b));
  sg_batch_destroy(b);
  sg_model_destroy(m);
  free(blob); free(mask); free(k); free(traj); free(status);
  return 0;
}
