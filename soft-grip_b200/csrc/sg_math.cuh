// sg_math.cuh -- small device math + narrowphase primitives shared by the kernels (sm_100a device code; also
// compiled for the host by the SIMT emulator in tests/simt, which is test infrastructure only).
#pragma once
#include <stdint.h>

namespace sg {

#define SG_MINVAL 1e-15
#define SG_MAXVAL 1e10
#define FULLMASK 0xffffffffu
// ---------------------------------------------------------------------------------------------
// small math
// ---------------------------------------------------------------------------------------------
template <typename T> __device__ __forceinline__ T tsqrt(T x);
template <> __device__ __forceinline__ float tsqrt<float>(float x) { return sqrtf(x); }
template <> __device__ __forceinline__ double tsqrt<double>(double x) { return sqrt(x); }
template <typename T> __device__ __forceinline__ T tabs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T tmin(T a, T b) { return a < b ? a : b; }
template <typename T> __device__ __forceinline__ T tmax(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ T tpow(T x, T y);
template <> __device__ __forceinline__ float tpow<float>(float x, float y) { return powf(x, y); }
template <> __device__ __forceinline__ double tpow<double>(double x, double y) { return pow(x, y); }
template <typename T> __device__ __forceinline__ void tsincos(T x, T* s, T* c);
template <> __device__ __forceinline__ void tsincos<float>(float x, float* s, float* c) { sincosf(x, s, c); }
template <> __device__ __forceinline__ void tsincos<double>(double x, double* s, double* c) { sincos(x, s, c); }

template <typename T> __device__ __forceinline__ T dot3(const T* a, const T* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
template <typename T> __device__ __forceinline__ void cross3(T* r, const T* a, const T* b) {
  T x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ T normalize3(T* a) {
  T n = tsqrt(dot3(a, a));
  if (n < T(SG_MINVAL)) { a[0] = 1; a[1] = 0; a[2] = 0; } else { T i = T(1) / n; a[0] *= i; a[1] *= i; a[2] *= i; }
  return n;
}
template <typename T> __device__ __forceinline__ void matvec3(T* r, const T* R, const T* v) {
  T x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2], z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void matTvec3(T* r, const T* R, const T* v) {
  T x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2], z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
template <typename T> __device__ __forceinline__ void matmul3(T* C, const T* A, const T* B) {
  T t[9];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
#pragma unroll
  for (int i = 0; i < 9; i++) C[i] = t[i];
}
template <typename T> __device__ __forceinline__ T warp_sum(T x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(FULLMASK, x, o);
  return x;
}
__device__ __forceinline__ int warp_max_i(int x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { int y = __shfl_xor_sync(FULLMASK, x, o); x = x > y ? x : y; }
  return x;
}

// getimpedance with pre-sanitised solimp (SURVEY App. A1 "Impedance per row")
template <typename T> __device__ __forceinline__ T impedance(const double* si, T pos) {
  T d0 = T(si[0]), d1 = T(si[1]), w = T(si[2]), mid = T(si[3]), pw = T(si[4]);
  if (d0 == d1 || w <= T(SG_MINVAL)) return T(0.5) * (d0 + d1);
  T x = tabs(pos / w);
  if (x >= T(1)) return d1;
  if (x <= T(0)) return d0;
  T y;
  if (pw == T(1)) y = x;
  else if (x <= mid) y = tpow(x, pw) / tpow(mid, pw - T(1));
  else y = T(1) - tpow(T(1) - x, pw) / tpow(T(1) - mid, pw - T(1));
  return d0 + y * (d1 - d0);
}

// ---------------------------------------------------------------------------------------------
// narrowphase
// ---------------------------------------------------------------------------------------------
template <typename T> struct RawCon { T dist, pos[3], nrm[3], hint[3]; };

// mjraw_SphereBox: normal from the sphere towards the box
template <typename T>
__device__ int sphere_box(RawCon<T>& con, const T* spos, T radius, const T* bpos, const T* bmat, const T* bsize) {
  T tmp[3], center[3], clamped[3], pos[3], nrm[3];
#pragma unroll
  for (int k = 0; k < 3; k++) tmp[k] = spos[k] - bpos[k];
  matTvec3(center, bmat, tmp);
#pragma unroll
  for (int k = 0; k < 3; k++) { clamped[k] = tmax(-bsize[k], tmin(bsize[k], center[k])); nrm[k] = clamped[k] - center[k]; }
  T dist = tsqrt(dot3(nrm, nrm));
  if (dist - radius > T(0)) return 0;
  if (dist <= T(SG_MINVAL)) {
    T closest = T(2) * (bsize[0] + bsize[1] + bsize[2]); int kbest = 0;
#pragma unroll
    for (int i = 0; i < 6; i++) {
      T fd = tabs(((i & 1) ? T(1) : T(-1)) * bsize[i >> 1] - center[i >> 1]);
      if (fd < closest) { closest = fd; kbest = i; }
    }
    nrm[0] = nrm[1] = nrm[2] = 0;
    T sg = (kbest & 1) ? T(-1) : T(1);
    if ((kbest >> 1) == 0) nrm[0] = sg; else if ((kbest >> 1) == 1) nrm[1] = sg; else nrm[2] = sg;
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = center[k] + nrm[k] * (radius - closest) / T(2);
    con.dist = -closest - radius;
  } else {
#pragma unroll
    for (int k = 0; k < 3; k++) nrm[k] /= dist;
#pragma unroll
    for (int k = 0; k < 3; k++) pos[k] = T(0.5) * (clamped[k] + center[k] + nrm[k] * radius);
    con.dist = dist - radius;
  }
  matvec3(con.nrm, bmat, nrm);
  matvec3(tmp, bmat, pos);
#pragma unroll
  for (int k = 0; k < 3; k++) { con.pos[k] = tmp[k] + bpos[k]; con.hint[k] = 0; }
  return 1;
}

template <typename T>
__device__ __forceinline__ T seg_box_grad(const T* c, const T* h, const T* s, T t, T* d2) {
  T g = 0, q = 0;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    T p = c[k] + t * h[k];
    T e = p - tmax(-s[k], tmin(s[k], p));
    g += h[k] * e; q += e * e;
  }
  if (d2) *d2 = q;
  return g;
}

// capsule-box as defined in oracle/sg_oracle.c (closest point of the segment by the convex signed
// distance, then a second sphere test at the far end)
template <typename T>
__device__ int capsule_box(RawCon<T>* con, const T* cpos, const T* axis_w, T radius, T hl, const T* bpos, const T* bmat, const T* bsize) {
  T tmp[3], c[3], ax[3], h[3];
#pragma unroll
  for (int k = 0; k < 3; k++) tmp[k] = cpos[k] - bpos[k];
  matTvec3(c, bmat, tmp);
  matTvec3(ax, bmat, axis_w);
#pragma unroll
  for (int k = 0; k < 3; k++) h[k] = ax[k] * hl;
  T tlo = -1, thi = 1, d2;
  T glo = seg_box_grad(c, h, bsize, T(-1), (T*)nullptr), ghi = seg_box_grad(c, h, bsize, T(1), (T*)nullptr);
  T tstar;
  if (glo >= T(0)) tstar = -1;
  else if (ghi <= T(0)) tstar = 1;
  else {
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (tabs(h[k]) < T(SG_MINVAL)) continue;
#pragma unroll
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        T t = (T(sgn) * bsize[k] - c[k]) / h[k];
        if (t <= tlo || t >= thi) continue;
        T g = seg_box_grad(c, h, bsize, t, (T*)nullptr);
        if (g <= T(0)) { tlo = t; glo = g; } else { thi = t; ghi = g; }
      }
    }
    tstar = (ghi - glo > T(SG_MINVAL)) ? tlo + (T(0) - glo) * (thi - tlo) / (ghi - glo) : tlo;
  }
  seg_box_grad(c, h, bsize, tstar, &d2);
  if (d2 <= T(SG_MINVAL) * T(SG_MINVAL)) {
    T t0 = -1, t1 = 1;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      if (tabs(h[k]) < T(SG_MINVAL)) continue;
      T ta = (-bsize[k] - c[k]) / h[k], tb = (bsize[k] - c[k]) / h[k];
      if (ta > tb) { T x = ta; ta = tb; tb = x; }
      if (ta > t0) t0 = ta;
      if (tb < t1) t1 = tb;
    }
    const bool in0 = (t0 <= T(-1)), in1 = (t1 >= T(1));
    if (in0 && in1) {
      T dm = T(1e30), dp = T(1e30);
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const T a0 = bsize[k] - c[k], a1 = bsize[k] + c[k];
        dm = tmin(dm, tmin(a0 + h[k], a1 - h[k]));     // depth at t = -1
        dp = tmin(dp, tmin(a0 - h[k], a1 + h[k]));     // depth at t = +1
      }
      tstar = (dp > dm) ? T(1) : T(-1);
    } else if (in0) tstar = T(-1);
    else if (in1) tstar = T(1);
    else tstar = T(0.5) * (t0 + t1);
  }
  int n = 0; T sp[3];
#pragma unroll
  for (int k = 0; k < 3; k++) sp[k] = cpos[k] + axis_w[k] * (tstar * hl);
  n += sphere_box(con[n], sp, radius, bpos, bmat, bsize);
  T t2 = (tstar >= T(0)) ? T(-1) : T(1);
#pragma unroll
  for (int k = 0; k < 3; k++) sp[k] = cpos[k] + axis_w[k] * (t2 * hl);
  n += sphere_box(con[n], sp, radius, bpos, bmat, bsize);
  return n;
}

// mjc_PlaneCapsule
template <typename T>
__device__ int plane_capsule(RawCon<T>* con, const T* ppos, const T* pmat, const T* cpos, const T* axis, T radius, T hl) {
  T nrm[3] = {pmat[2], pmat[5], pmat[8]};
  int n = 0;
#pragma unroll
  for (int side = 1; side >= -1; side -= 2) {
    T sp[3], dif[3];
#pragma unroll
    for (int k = 0; k < 3; k++) { sp[k] = cpos[k] + T(side) * axis[k] * hl; dif[k] = sp[k] - ppos[k]; }
    T dist = dot3(dif, nrm) - radius;
    if (dist > T(0)) continue;
    con[n].dist = dist;
#pragma unroll
    for (int k = 0; k < 3; k++) { con[n].pos[k] = sp[k] - nrm[k] * (radius + T(0.5) * dist); con[n].nrm[k] = nrm[k]; con[n].hint[k] = axis[k]; }
    n++;
  }
  return n;
}

template <typename T>
__device__ bool box_box_overlap(const T* p1, const T* R1, const T* s1, const T* p2, const T* R2, const T* s2) {
  T Rr[9], A[9], T3[3], tmp[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) { Rr[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j]; A[3 * i + j] = tabs(Rr[3 * i + j]) + T(1e-12); }
#pragma unroll
  for (int k = 0; k < 3; k++) tmp[k] = p2[k] - p1[k];
  matTvec3(T3, R1, tmp);
  for (int i = 0; i < 3; i++) if (tabs(T3[i]) > s1[i] + s2[0] * A[3 * i] + s2[1] * A[3 * i + 1] + s2[2] * A[3 * i + 2]) return false;
  for (int j = 0; j < 3; j++) if (tabs(T3[0] * Rr[j] + T3[1] * Rr[3 + j] + T3[2] * Rr[6 + j]) > s2[j] + s1[0] * A[j] + s1[1] * A[3 + j] + s1[2] * A[6 + j]) return false;
  for (int i = 0; i < 3; i++)
    for (int j = 0; j < 3; j++) {
      int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
      T ra = s1[i1] * A[3 * i2 + j] + s1[i2] * A[3 * i1 + j];
      T rb = s2[j1] * A[3 * i + j2] + s2[j2] * A[3 * i + j1];
      if (tabs(T3[i2] * Rr[3 * i1 + j] - T3[i1] * Rr[3 * i2 + j]) > ra + rb) return false;
    }
  return true;
}

// mju_makeFrame: f[0..2] normal, f[3..5] hint -> orthonormal frame
template <typename T> __device__ void make_frame(T* f) {
  normalize3(f);
  if (tsqrt(dot3(f + 3, f + 3)) < T(0.5)) { f[3] = f[4] = f[5] = 0; if (f[1] < T(0.5) && f[1] > T(-0.5)) f[4] = 1; else f[5] = 1; }
  T dp = dot3(f, f + 3);
#pragma unroll
  for (int k = 0; k < 3; k++) f[3 + k] -= f[k] * dp;
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

// mju_QCQP2
template <typename T> __device__ int qcqp2(T* res, T A11i, T A12i, T A22i, const T* bin, T d0, T d1, T r) {
  T b1 = bin[0] * d0, b2 = bin[1] * d1;
  T A11 = A11i * d0 * d0, A22 = A22i * d1 * d1, A12 = A12i * d0 * d1;
  T la = 0, v1 = 0, v2 = 0;
  for (int iter = 0; iter < 20; iter++) {
    T det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < T(1e-10)) { res[0] = 0; res[1] = 0; return 0; }
    T detinv = T(1) / det, P11 = (A22 + la) * detinv, P22 = (A11 + la) * detinv, P12 = -A12 * detinv;
    v1 = -P11 * b1 - P12 * b2; v2 = -P12 * b1 - P22 * b2;
    T val = v1 * v1 + v2 * v2 - r * r;
    if (val < T(1e-10)) break;
    T deriv = T(-2) * (P11 * v1 * v1 + T(2) * P12 * v1 * v2 + P22 * v2 * v2);
    T delta = -val / deriv;
    if (delta < T(1e-10)) break;
    la += delta;
  }
  res[0] = v1 * d0; res[1] = v2 * d1;
  return la != T(0);
}

}  // namespace sg
