// sg_kernels.cuh -- sm_100a device code of the batched soft-gripper simulator.
//
// One warp (= one CTA of 32 threads) owns one world.  The whole mj_step of that world
// (ref call site: environment/manenv.py:49 `self.env.step()`) runs out
// of the CTA's shared memory:
//
//   gripper()        finger-chain kinematics, inertia, bias, actuation, sensor pre-data   (1 lane / chain)
//   collide()        broadphase (lane / pair) -> candidate list -> narrowphase (lane / candidate),
//                    contacts emitted in MuJoCo's canonical order by ballot compaction
//   rows_setup()     equality / tendon / limit rows: pos, impedance, R, aref                  (lane / row)
//   smooth()         slider spring-dampers, volume-tendon spring-damper, gravity             (lane / dof)
//   warmstart()      forces from qacc_warmstart, dual cost test, qacc = qacc_smooth + M^-1 J^T f
//   pgs()            projected Gauss-Seidel in exact MuJoCo row order: the equality block is swept by
//                    *dependency levels* (rows of one level touch disjoint dofs, so lanes update them
//                    concurrently with results identical to the sequential sweep), then the dense
//                    tendon row (warp-shuffle reduction), limits, elliptic contact blocks by chain level
//   finish()         qacc from the final forces, accelerometers, implicit-damping Euler, NaN checks
//
// No tensor cores: nothing here is a dense contraction (largest dense objects: 4x4 finger inertia
// blocks, 3x3 contact blocks).  The kernel is bound by the dependent-issue latency of the Gauss-Seidel
// sweep, so throughput comes from many resident worlds per SM (small shared-memory footprint).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sg_plan.hpp"
#include "sg_math.cuh"

namespace sg {

#ifndef SG_ST_CON_FULL_BIT
#define SG_ST_CON_FULL_BIT 2
#define SG_ST_UNSUPPORTED_BIT 8
#endif

// shared-memory layout of one world; offsets in elements of T (then ints)
struct SmemLayout {
  int q, v, a, qs;                 // nv each (qs: qacc_smooth for finger dofs, qfrc_smooth for sliders)
  int jtf;                         // nv  J^T f scratch
  int rf, raref, rR;               // nrow each, schedule order
  int ten;                         // 8: f, aref, R, A, L, Ldot, Ft, pad
  int lim;                         // 3*nfd: f, aref, R
  int g_axis, g_anchor;            // 3*nfd each
  int g_minv;                      // 16*nchain
  int g_box;                       // 12 per moving box (MAXCHAIN*MAXCB)
  int g_actdot;                    // nu
  int s_jv, s_ab, s_rot;           // per sensor: 12, 3, 9
  int sens;                        // nsd
  int c_jg, c_ns, c_aref, c_R, c_A, c_f;   // per contact: 12, 3, 3, 2, 6, 3
  int nT;
  int i_lim;                       // 2*nfd : dof, sign
  int i_con;                       // 3*maxcon : chain, slider, level
  int i_cand;                      // maxcand
  int i_misc;                      // 8: ncon, nlim, maxlev, status, touch, ncon_total, iters, ncand
  int nI;
  int bytes;
};

template <typename T>
inline SmemLayout make_layout(const PlanDims& D) {
  SmemLayout L{};
  int o = 0;
  auto take = [&](int n) { int r = o; o += n; return r; };
  L.q = take(D.nv); L.v = take(D.nv); L.a = take(D.nv); L.qs = take(D.nv); L.jtf = take(D.nv);
  L.rf = take(D.nrow); L.raref = take(D.nrow); L.rR = take(D.nrow);
  L.ten = take(8); L.lim = take(3 * MAXFD);
  L.g_axis = take(3 * MAXFD); L.g_anchor = take(3 * MAXFD); L.g_minv = take(16 * MAXCHAIN);
  L.g_box = take(12 * MAXCHAIN * MAXCB); L.g_actdot = take(D.nu > 0 ? D.nu : 1);
  L.s_jv = take(12 * MAXSENS); L.s_ab = take(3 * MAXSENS); L.s_rot = take(9 * MAXSENS); L.sens = take(D.nsd > 0 ? D.nsd : 1);
  L.c_jg = take(12 * D.maxcon); L.c_ns = take(3 * D.maxcon); L.c_aref = take(3 * D.maxcon); L.c_R = take(2 * D.maxcon);
  L.c_A = take(6 * D.maxcon); L.c_f = take(3 * D.maxcon);
  if (o & 1) o++;
  L.nT = o;
  int io = 0;
  auto takei = [&](int n) { int r = io; io += n; return r; };
  L.i_lim = takei(2 * MAXFD); L.i_con = takei(3 * D.maxcon); L.i_cand = takei(D.maxcand); L.i_misc = takei(8);
  L.nI = io;
  L.bytes = (int)(sizeof(T) * L.nT + sizeof(int) * L.nI);
  L.bytes = (L.bytes + 15) & ~15;
  return L;
}

enum { MI_NCON = 0, MI_NLIM = 1, MI_MAXLEV = 2, MI_STATUS = 3, MI_TOUCH = 4, MI_NCONTOT = 5, MI_ITERS = 6, MI_NCAND = 7 };

template <typename T>
struct KArgs {
  PlanDims D;
  SmemLayout L;
  const T* tab;
  const int* itab;
  int nworlds;
  // state, world-major
  T *qpos, *qvel, *warm, *act, *ctrl;
  // per-world parameters (may be null)
  const double *p_stiff, *p_damp, *p_tdamp, *p_objoff;
  int* status;          // [W] accumulated bits
  // outputs
  T* sens_out;          // step: [W][nsd]; rollout: [W][nrows][nsd]
  int* touch_out;       // step: [W];      rollout: [W][nrows]
  // step mode
  int nsub;
  int integrate;        // 0: mj_forward only
  // rollout mode
  int rollout, sim_start, sim_step, nrows;
  const int* ctrl_event;     // dev [nrows]
  const double* ctrl_value;  // dev [nrows][nu]
  // diagnostics
  int debug_world;
  double* debug_out;    // layout documented in sg_api.cu
  int debug_cap;
};

// ---------------------------------------------------------------------------------------------
// the world
// ---------------------------------------------------------------------------------------------
template <typename T>
struct World {
  const KArgs<T>& K;
  const PlanDims& D;
  const SmemLayout& L;
  T* sm;
  int* smi;
  int lane, w;
  T kw, dw, tdw, off[3];       // per-world parameter overrides (kw<0 etc. = none)
  T sl_iw0;                    // uniform inverse slider mass is NOT assumed; this is only a cache for lane-local use

  __device__ World(const KArgs<T>& k, T* s, int wid) : K(k), D(k.D), L(k.L), sm(s), smi((int*)(s + k.L.nT)), lane(threadIdx.x & 31), w(wid) {}

  __device__ __forceinline__ const T* tab(int o) const { return K.tab + o; }
  __device__ __forceinline__ const int* itab(int o) const { return K.itab + o; }
  __device__ __forceinline__ T* q() { return sm + L.q; }
  __device__ __forceinline__ T* v() { return sm + L.v; }
  __device__ __forceinline__ T* a() { return sm + L.a; }
  __device__ __forceinline__ T* qs() { return sm + L.qs; }
  __device__ __forceinline__ int& misc(int i) { return smi[L.i_misc + i]; }

  __device__ void load_params() {
    kw = K.p_stiff ? T(K.p_stiff[w]) : T(-1);
    dw = K.p_damp ? T(K.p_damp[w]) : T(-1);
    tdw = K.p_tdamp ? T(K.p_tdamp[w]) : T(-1);
#pragma unroll
    for (int k = 0; k < 3; k++) off[k] = T(D.obj_pos[k]) + (K.p_objoff ? T(K.p_objoff[3 * (size_t)w + k]) : T(0));
  }
  __device__ __forceinline__ T stiffness(int e) const { return (kw >= T(0) && itab(D.io_kmask)[e]) ? kw : tab(D.o_sl_k0)[e]; }
  __device__ __forceinline__ T damping(int e) const { return dw >= T(0) ? dw : tab(D.o_sl_d0)[e]; }
  __device__ __forceinline__ T ten_stiffness() const { return (kw >= T(0) && D.stiff_tendon0) ? kw : T(D.ten_k0); }
  __device__ __forceinline__ T ten_damping() const { return tdw >= T(0) ? tdw : T(D.ten_d0); }

  // mj_resetData
  __device__ void reset_state() {
    for (int i = lane; i < D.nv; i += 32) { q()[i] = 0; v()[i] = 0; a()[i] = 0; }
    if (lane < D.nu) { K.act[(size_t)w * D.nu + lane] = 0; K.ctrl[(size_t)w * D.nu + lane] = 0; }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // finger chains: kinematics, inertia, bias, actuation, sensors (one lane per chain)
  // world-frame Jacobian formulation (independent of the oracle's com-based spatial algebra)
  // ------------------------------------------------------------------------------------------
  __device__ void gripper(int c) {
    const T* ch = tab(D.o_chain + c * CH_STRIDE);
    const int nd = D.ncd[c], nb = D.ncb[c], dof0 = D.chain_dof0[c];
    T P[3], R[9];
#pragma unroll
    for (int k = 0; k < 3; k++) P[k] = ch[CH_BASEPOS + k];
#pragma unroll
    for (int k = 0; k < 9; k++) R[k] = ch[CH_BASEROT + k];
    T axis[MAXCD][3], anch[MAXCD][3], bpos[MAXCB][3], brot[MAXCB][9], com[MAXCB][3], Iw[MAXCB][6], om[MAXCB][3];
    int nsup[MAXCB];
    int j = 0;
#pragma unroll
    for (int k = 0; k < MAXCB; k++) {
      if (k >= nb) break;
      const T* cb = ch + CH_BODY + k * CB_STRIDE;
      T pos[3], Rc[9], t[3];
      matvec3(t, R, cb + CB_POS);
#pragma unroll
      for (int i = 0; i < 3; i++) pos[i] = P[i] + t[i];
      matmul3(Rc, R, cb + CB_ROT);
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) {
        if (jj != j || j >= nd) continue;
        const T* cd = ch + CH_DOF + j * CD_STRIDE;
        if ((int)cd[CD_BODY] != k) continue;
        matvec3(axis[j], Rc, cd + CD_AXIS);
        matvec3(t, Rc, cd + CD_JPOS);
#pragma unroll
        for (int i = 0; i < 3; i++) anch[j][i] = pos[i] + t[i];
        // Rodrigues rotation about the local axis by q
        T s, co; tsincos(q()[dof0 + j], &s, &co);
        const T ux = cd[CD_AXIS], uy = cd[CD_AXIS + 1], uz = cd[CD_AXIS + 2], oc = T(1) - co;
        T Rj[9] = {co + ux * ux * oc, ux * uy * oc - uz * s, ux * uz * oc + uy * s,
                   uy * ux * oc + uz * s, co + uy * uy * oc, uy * uz * oc - ux * s,
                   uz * ux * oc - uy * s, uz * uy * oc + ux * s, co + uz * uz * oc};
        matmul3(Rc, Rc, Rj);
        matvec3(t, Rc, cd + CD_JPOS);
#pragma unroll
        for (int i = 0; i < 3; i++) pos[i] = anch[j][i] - t[i];
        j++;
      }
      nsup[k] = j;
#pragma unroll
      for (int i = 0; i < 3; i++) { bpos[k][i] = pos[i]; P[i] = pos[i]; }
#pragma unroll
      for (int i = 0; i < 9; i++) { brot[k][i] = Rc[i]; R[i] = Rc[i]; }
      matvec3(t, Rc, cb + CB_IPOS);
#pragma unroll
      for (int i = 0; i < 3; i++) com[k][i] = pos[i] + t[i];
      T Ri[9]; matmul3(Ri, Rc, cb + CB_IROT);
      const T I0 = cb[CB_INERTIA], I1 = cb[CB_INERTIA + 1], I2 = cb[CB_INERTIA + 2];
      Iw[k][0] = Ri[0] * Ri[0] * I0 + Ri[1] * Ri[1] * I1 + Ri[2] * Ri[2] * I2;   // xx
      Iw[k][1] = Ri[3] * Ri[3] * I0 + Ri[4] * Ri[4] * I1 + Ri[5] * Ri[5] * I2;   // yy
      Iw[k][2] = Ri[6] * Ri[6] * I0 + Ri[7] * Ri[7] * I1 + Ri[8] * Ri[8] * I2;   // zz
      Iw[k][3] = Ri[0] * Ri[3] * I0 + Ri[1] * Ri[4] * I1 + Ri[2] * Ri[5] * I2;   // xy
      Iw[k][4] = Ri[0] * Ri[6] * I0 + Ri[1] * Ri[7] * I1 + Ri[2] * Ri[8] * I2;   // xz
      Iw[k][5] = Ri[3] * Ri[6] * I0 + Ri[4] * Ri[7] * I1 + Ri[5] * Ri[8] * I2;   // yz
      // box geom pose -> shared
      T* gb = sm + L.g_box + 12 * (c * MAXCB + k);
      matvec3(t, Rc, cb + CB_GPOS);
#pragma unroll
      for (int i = 0; i < 3; i++) gb[i] = pos[i] + t[i];
      T Rg[9]; matmul3(Rg, Rc, cb + CB_GROT);
#pragma unroll
      for (int i = 0; i < 9; i++) gb[3 + i] = Rg[i];
    }
#pragma unroll
    for (int jj = 0; jj < MAXCD; jj++) {
      if (jj >= nd) break;
#pragma unroll
      for (int i = 0; i < 3; i++) { sm[L.g_axis + 3 * (dof0 + jj) + i] = axis[jj][i]; sm[L.g_anchor + 3 * (dof0 + jj) + i] = anch[jj][i]; }
    }
    // velocity-dependent terms
    T qv[MAXCD], da[MAXCD][3], va[MAXCD][3];
#pragma unroll
    for (int jj = 0; jj < MAXCD; jj++) qv[jj] = jj < nd ? v()[dof0 + jj] : T(0);
#pragma unroll
    for (int jj = 0; jj < MAXCD; jj++) {
      T wb[3] = {0, 0, 0};
      va[jj][0] = va[jj][1] = va[jj][2] = 0;
      if (jj < nd) {
#pragma unroll
        for (int i = 0; i < MAXCD; i++) {
          if (i >= jj) break;
          T r[3] = {anch[jj][0] - anch[i][0], anch[jj][1] - anch[i][1], anch[jj][2] - anch[i][2]}, t[3];
          cross3(t, axis[i], r);
#pragma unroll
          for (int k = 0; k < 3; k++) { wb[k] += axis[i][k] * qv[i]; va[jj][k] += t[k] * qv[i]; }
        }
        cross3(da[jj], wb, axis[jj]);
      } else { da[jj][0] = da[jj][1] = da[jj][2] = 0; }
    }
    auto point_terms = [&](const T* p, int ns_, T Jv[MAXCD][3], T* vp, T* ab) {
      vp[0] = vp[1] = vp[2] = 0; ab[0] = ab[1] = ab[2] = 0;
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) {
        if (jj >= ns_) { Jv[jj][0] = Jv[jj][1] = Jv[jj][2] = 0; continue; }
        T r[3] = {p[0] - anch[jj][0], p[1] - anch[jj][1], p[2] - anch[jj][2]};
        cross3(Jv[jj], axis[jj], r);
#pragma unroll
        for (int k = 0; k < 3; k++) vp[k] += Jv[jj][k] * qv[jj];
      }
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) {
        if (jj >= ns_) continue;
        T r[3] = {p[0] - anch[jj][0], p[1] - anch[jj][1], p[2] - anch[jj][2]}, t1[3], t2[3];
        T dv[3] = {vp[0] - va[jj][0], vp[1] - va[jj][1], vp[2] - va[jj][2]};
        cross3(t1, da[jj], r); cross3(t2, axis[jj], dv);
#pragma unroll
        for (int k = 0; k < 3; k++) ab[k] += (t1[k] + t2[k]) * qv[jj];
      }
    };
    T M[MAXCD][MAXCD], frc[MAXCD];
#pragma unroll
    for (int i = 0; i < MAXCD; i++) { frc[i] = 0;
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) M[i][jj] = (i == jj && i >= nd) ? T(1) : T(0); }
    const T g[3] = {T(D.g[0]), T(D.g[1]), T(D.g[2])};
#pragma unroll
    for (int k = 0; k < MAXCB; k++) {
      if (k >= nb) break;
      const T* cb = ch + CH_BODY + k * CB_STRIDE;
      const T mass = cb[CB_MASS];
      T Jv[MAXCD][3], vc[3], ac[3], al[3] = {0, 0, 0};
      om[k][0] = om[k][1] = om[k][2] = 0;
      point_terms(com[k], nsup[k], Jv, vc, ac);
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) {
        if (jj >= nsup[k]) break;
#pragma unroll
        for (int i = 0; i < 3; i++) { om[k][i] += axis[jj][i] * qv[jj]; al[i] += da[jj][i] * qv[jj]; }
      }
      auto Imul = [&](const T* x, T* y) {
        y[0] = Iw[k][0] * x[0] + Iw[k][3] * x[1] + Iw[k][4] * x[2];
        y[1] = Iw[k][3] * x[0] + Iw[k][1] * x[1] + Iw[k][5] * x[2];
        y[2] = Iw[k][4] * x[0] + Iw[k][5] * x[1] + Iw[k][2] * x[2];
      };
      T F[3] = {mass * (ac[0] - g[0]), mass * (ac[1] - g[1]), mass * (ac[2] - g[2])};
      T Ial[3], Iom[3], N[3];
      Imul(al, Ial); Imul(om[k], Iom); cross3(N, om[k], Iom);
#pragma unroll
      for (int i = 0; i < 3; i++) N[i] += Ial[i];
#pragma unroll
      for (int i = 0; i < MAXCD; i++) {
        if (i >= nsup[k]) break;
        frc[i] -= dot3(Jv[i], F) + dot3(axis[i], N);       // -qfrc_bias
        T Ia[3]; Imul(axis[i], Ia);
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++) {
          if (jj > i) break;
          T mij = mass * dot3(Jv[i], Jv[jj]) + dot3(axis[jj], Ia);
          M[i][jj] += mij;
          if (jj != i) M[jj][i] += mij;
        }
      }
    }
    // actuation through the chain's spatial tendon (cylinder: filter dynamics, force = gain*act)
    const T* ct = ch + CH_TEN;
    if (ct[CT_HAS] != T(0)) {
      const int kb = (int)ct[CT_BODY], u = (int)ct[CT_ACT];
      T s1[3], t[3], dir[3];
      matvec3(t, brot[kb], ct + CT_S1);
#pragma unroll
      for (int i = 0; i < 3; i++) { s1[i] = bpos[kb][i] + t[i]; dir[i] = s1[i] - ct[CT_S0 + i]; }
      normalize3(dir);
      if (u >= 0) {
        const T actv = K.act[(size_t)w * D.nu + u], ctrlv = K.ctrl[(size_t)w * D.nu + u];
        sm[L.g_actdot + u] = (ctrlv - actv) / tmax(T(SG_MINVAL), ct[CT_TIMECONST]);
        const T force = ct[CT_GAIN] * actv;
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++) {
          if (jj >= nsup[kb]) break;
          T r[3] = {s1[0] - anch[jj][0], s1[1] - anch[jj][1], s1[2] - anch[jj][2]}, jc[3];
          cross3(jc, axis[jj], r);
          frc[jj] += ct[CT_GEAR] * dot3(dir, jc) * force;
        }
      }
    }
    // M^-1 by Gauss-Jordan on the (padded) 4x4 SPD block
    T Mi[MAXCD][MAXCD];
#pragma unroll
    for (int i = 0; i < MAXCD; i++)
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) Mi[i][jj] = i == jj ? T(1) : T(0);
#pragma unroll
    for (int p = 0; p < MAXCD; p++) {
      const T ip = T(1) / M[p][p];
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) { M[p][jj] *= ip; Mi[p][jj] *= ip; }
#pragma unroll
      for (int i = 0; i < MAXCD; i++) {
        if (i == p) continue;
        const T f = M[i][p];
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++) { M[i][jj] -= f * M[p][jj]; Mi[i][jj] -= f * Mi[p][jj]; }
      }
    }
#pragma unroll
    for (int i = 0; i < MAXCD; i++) {
      T s = 0;
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) { sm[L.g_minv + 16 * c + 4 * i + jj] = Mi[i][jj]; s += Mi[i][jj] * frc[jj]; }
      if (i < nd) qs()[dof0 + i] = s;
    }
    // sensors on this chain (gyro now; accelerometer pre-data, finished after the solve)
    for (int s = 0; s < D.nsens; s++) {
      const T* se = tab(D.o_sens + s * SE_STRIDE);
      if ((int)se[SE_CHAIN] != c) continue;
      const int k = (int)se[SE_BODY], adr = (int)se[SE_ADR];
      T Rs[9]; matmul3(Rs, brot[k], se + SE_ROT);
      if ((int)se[SE_TYPE] == SENS_GYRO) {
        T o[3]; matTvec3(o, Rs, om[k]);
#pragma unroll
        for (int i = 0; i < 3; i++) sm[L.sens + adr + i] = o[i];
      } else {
        T p[3], t[3], Jv[MAXCD][3], vp[3], ab[3];
        matvec3(t, brot[k], se + SE_POS);
#pragma unroll
        for (int i = 0; i < 3; i++) p[i] = bpos[k][i] + t[i];
        point_terms(p, nsup[k], Jv, vp, ab);
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++)
#pragma unroll
          for (int i = 0; i < 3; i++) sm[L.s_jv + 12 * s + 3 * jj + i] = Jv[jj][i];
#pragma unroll
        for (int i = 0; i < 3; i++) sm[L.s_ab + 3 * s + i] = ab[i] - g[i];
#pragma unroll
        for (int i = 0; i < 9; i++) sm[L.s_rot + 9 * s + i] = Rs[i];
      }
    }
  }

  // ------------------------------------------------------------------------------------------
  // collision + contact rows
  // ------------------------------------------------------------------------------------------
  __device__ __forceinline__ void capsule_center(int e, T* c) {
    const T* c0 = tab(D.o_sl_cap0) + 3 * e; const T* ax = tab(D.o_sl_axis) + 3 * e;
    const T qe = q()[D.nfd + e];
#pragma unroll
    for (int k = 0; k < 3; k++) c[k] = off[k] + c0[k] + ax[k] * qe;
  }
  __device__ __forceinline__ void collider_pose(int ci, T* pos, T* rot) {
    const T* co = tab(D.o_coll + ci * CO_STRIDE);
    const int c = (int)co[CO_CHAIN];
    if (c < 0) {
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = co[CO_POS + k];
#pragma unroll
      for (int k = 0; k < 9; k++) rot[k] = co[CO_ROT + k];
    } else {
      const T* gb = sm + L.g_box + 12 * (c * MAXCB + (int)co[CO_BODY]);
#pragma unroll
      for (int k = 0; k < 3; k++) pos[k] = gb[k];
#pragma unroll
      for (int k = 0; k < 9; k++) rot[k] = gb[3 + k];
    }
  }

  // builds the three rows of one contact into slot `slot` (mj_instantiateContact + mj_makeImpedance +
  // mj_referenceConstraint + the diagonal block of efc_AR)
  __device__ void contact_rows(int slot, const RawCon<T>& rc, int ci, int e, T slider_sign, bool dbg, int dbg_index) {
    const T* co = tab(D.o_coll + ci * CO_STRIDE);
    const int c = (int)co[CO_CHAIN];
    T fr[9];
#pragma unroll
    for (int k = 0; k < 3; k++) { fr[k] = rc.nrm[k]; fr[3 + k] = rc.hint[k]; }
    make_frame(fr);
    T Jg[3][MAXCD];
    T* jg = sm + L.c_jg + 12 * slot;
    int nsupp = 0, dof0 = 0;
    if (c >= 0) {
      dof0 = D.chain_dof0[c];
      const T* ch = tab(D.o_chain + c * CH_STRIDE);
      const int kb = (int)co[CO_BODY];
      for (int jj = 0; jj < D.ncd[c]; jj++) if ((int)ch[CH_DOF + jj * CD_STRIDE + CD_BODY] <= kb) nsupp = jj + 1;
    }
#pragma unroll
    for (int jj = 0; jj < MAXCD; jj++) {
      T col[3] = {0, 0, 0};
      if (jj < nsupp) {
        const T* ax = sm + L.g_axis + 3 * (dof0 + jj); const T* an = sm + L.g_anchor + 3 * (dof0 + jj);
        T r[3] = {rc.pos[0] - an[0], rc.pos[1] - an[1], rc.pos[2] - an[2]};
        cross3(col, ax, r);
      }
#pragma unroll
      for (int r = 0; r < 3; r++) { Jg[r][jj] = dot3(fr + 3 * r, col); jg[4 * r + jj] = Jg[r][jj]; }
    }
    T ns[3] = {0, 0, 0}, iw_e = 0, biw = co[CO_BIW], ve = 0;
    if (e >= 0) {
      const T* ax = tab(D.o_sl_axis) + 3 * e;
#pragma unroll
      for (int r = 0; r < 3; r++) ns[r] = slider_sign * dot3(fr + 3 * r, ax);
      iw_e = T(1) / tab(D.o_sl_m)[e];
      biw += tab(D.o_sl_biw)[e];
      ve = v()[D.nfd + e];
    }
#pragma unroll
    for (int r = 0; r < 3; r++) sm[L.c_ns + 3 * slot + r] = ns[r];
    const T imp = impedance<T>(D.con_solimp, rc.dist);
    const T R0 = tmax(T(SG_MINVAL), (T(1) - imp) * biw / imp);
    const T R1 = R0 / tmax(T(SG_MINVAL), T(D.impratio));
    sm[L.c_R + 2 * slot] = R0; sm[L.c_R + 2 * slot + 1] = R1;
    // velocity, aref
    T vel[3];
#pragma unroll
    for (int r = 0; r < 3; r++) {
      T s = ns[r] * ve;
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) if (jj < nsupp) s += Jg[r][jj] * v()[dof0 + jj];
      vel[r] = s;
    }
    sm[L.c_aref + 3 * slot] = -T(D.con_B) * vel[0] - T(D.con_K) * imp * rc.dist;
    sm[L.c_aref + 3 * slot + 1] = -T(D.con_B) * vel[1];
    sm[L.c_aref + 3 * slot + 2] = -T(D.con_B) * vel[2];
    // A = Jg Minv Jg' + ns ns'/m + diag(R)
    T MJ[3][MAXCD];
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int i = 0; i < MAXCD; i++) {
        T s = 0;
        if (c >= 0) {
#pragma unroll
          for (int jj = 0; jj < MAXCD; jj++) s += sm[L.g_minv + 16 * c + 4 * i + jj] * Jg[r][jj];
        }
        MJ[r][i] = s;
      }
    T* A = sm + L.c_A + 6 * slot;
    int idx = 0;
#pragma unroll
    for (int r = 0; r < 3; r++)
#pragma unroll
      for (int s2 = r; s2 < 3; s2++) {
        T s = ns[r] * ns[s2] * iw_e;
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++) s += Jg[r][jj] * MJ[s2][jj];
        if (r == s2) s += (r == 0 ? R0 : R1);
        A[idx++] = s;    // order: 00 01 02 11 12 22
      }
    smi[L.i_con + 3 * slot] = c; smi[L.i_con + 3 * slot + 1] = e; smi[L.i_con + 3 * slot + 2] = 0;
    if (dbg && K.debug_out) {
      double* o = K.debug_out + 64 + 16 * (size_t)dbg_index;   // dist, pos3, frame9
      if (64 + 16 * (dbg_index + 1) <= K.debug_cap) {
        o[0] = (double)rc.dist;
        for (int k = 0; k < 3; k++) o[1 + k] = (double)rc.pos[k];
        for (int k = 0; k < 9; k++) o[4 + k] = (double)fr[k];
      }
    }
  }

  __device__ void collide() {
    if (lane == 0) { misc(MI_NCON) = 0; misc(MI_TOUCH) = 0; misc(MI_NCONTOT) = 0; misc(MI_NCAND) = 0; }
    __syncwarp();
    const bool dbg = (w == K.debug_world);
    // ---- broadphase: bounding spheres, candidates compacted in pair order ----
    int ncand = 0, flags = 0;
    for (int base = 0; base < D.npair; base += 32) {
      const int p = base + lane;
      bool pass = false;
      if (p < D.npair) {
        const int pt = itab(D.io_pair_t)[p], pa = itab(D.io_pair_a)[p], pb = itab(D.io_pair_b)[p];
        const T* co = tab(D.o_coll + pa * CO_STRIDE);
        T c2[3], rb2;
        if (pt == PAIR_PLANE_CAPSULE || pt == PAIR_BOX_CAPSULE) { capsule_center(pb, c2); rb2 = T(D.cap_r + D.cap_hl); }
        else if (pt == PAIR_SPHERE_BOX) {
#pragma unroll
          for (int k = 0; k < 3; k++) c2[k] = off[k] + T(D.sph_pos[k]);
          rb2 = T(D.sph_r);
        } else { T rot[9]; collider_pose(pb, c2, rot); rb2 = tab(D.o_coll + pb * CO_STRIDE)[CO_RBOUND]; }
        T c1[3], rot1[9];
        collider_pose(pa, c1, rot1);
        T dif[3] = {c2[0] - c1[0], c2[1] - c1[1], c2[2] - c1[2]};
        if ((int)co[CO_TYPE] == GEOM_PLANE) { T nrm[3] = {rot1[2], rot1[5], rot1[8]}; pass = !(dot3(dif, nrm) > rb2); }
        else { T bound = co[CO_RBOUND] + rb2; pass = !(dot3(dif, dif) > bound * bound); }
        if (pass && pt == PAIR_BOX_CAPSULE) {
          // mid-phase (prunes only): the capsule's bounding box in the frame of the box must overlap the box
          T dl[3], al[3];
          matTvec3(dl, rot1, dif);
          matTvec3(al, rot1, tab(D.o_sl_axis) + 3 * pb);
#pragma unroll
          for (int k = 0; k < 3; k++) if (tabs(dl[k]) > co[CO_SIZE + k] + T(D.cap_r) + T(D.cap_hl) * tabs(al[k])) pass = false;
        }
      }
      const unsigned m = __ballot_sync(FULLMASK, pass);
      if (pass) {
        const int slot = ncand + __popc(m & ((1u << lane) - 1));
        if (slot < D.maxcand) smi[L.i_cand + slot] = p; else flags |= SG_ST_CON_FULL_BIT;
      }
      ncand += __popc(m);
    }
    if (ncand > D.maxcand) ncand = D.maxcand;
    __syncwarp();
    // ---- narrowphase over the candidate list; contacts keep the pair order ----
    int ncon = 0, ncontot = 0, touch = 0;
    for (int base = 0; base < ncand; base += 32) {
      const int ci_ = base + lane;
      RawCon<T> rc[2];
      int n = 0, pa = 0, e = -1; T ssign = 0;
      if (ci_ < ncand) {
        const int p = smi[L.i_cand + ci_];
        const int pt = itab(D.io_pair_t)[p]; pa = itab(D.io_pair_a)[p]; const int pb = itab(D.io_pair_b)[p];
        const T* co = tab(D.o_coll + pa * CO_STRIDE);
        T c1[3], rot1[9];
        collider_pose(pa, c1, rot1);
        T size1[3] = {co[CO_SIZE], co[CO_SIZE + 1], co[CO_SIZE + 2]};
        int mask = (int)co[CO_MASK];
        if (pt == PAIR_PLANE_CAPSULE) {
          T cc[3]; capsule_center(pb, cc);
          n = plane_capsule(rc, c1, rot1, cc, tab(D.o_sl_axis) + 3 * pb, T(D.cap_r), T(D.cap_hl));
          e = pb; ssign = 1; mask |= (int)D.cap_mask;
        } else if (pt == PAIR_BOX_CAPSULE) {
          T cc[3]; capsule_center(pb, cc);
          n = capsule_box(rc, cc, tab(D.o_sl_axis) + 3 * pb, T(D.cap_r), T(D.cap_hl), c1, rot1, size1);
          e = pb; ssign = -1; mask |= (int)D.cap_mask;
        } else if (pt == PAIR_SPHERE_BOX) {
          T sc[3];
#pragma unroll
          for (int k = 0; k < 3; k++) sc[k] = off[k] + T(D.sph_pos[k]);
          n = sphere_box(rc[0], sc, T(D.sph_r), c1, rot1, size1);
          e = -1; mask |= (int)D.sph_mask;
        } else if (pt == PAIR_BOX_BOX) {
          const T* co2 = tab(D.o_coll + pb * CO_STRIDE);
          T c2[3], rot2[9]; collider_pose(pb, c2, rot2);
          T size2[3] = {co2[CO_SIZE], co2[CO_SIZE + 1], co2[CO_SIZE + 2]};
          if (box_box_overlap(c1, rot1, size1, c2, rot2, size2)) flags |= SG_ST_UNSUPPORTED_BIT;
        } else flags |= SG_ST_UNSUPPORTED_BIT;
        if (n > 0) { touch |= (1 << 30); if (mask & 1) touch |= (mask >> 1); }
      }
      // contacts with dist >= 0 (== includemargin) exist but carry no constraint rows
      const int n_act = (n > 0 && rc[0].dist < T(0) ? 1 : 0) + (n > 1 && rc[1].dist < T(0) ? 1 : 0);
      const unsigned m1 = __ballot_sync(FULLMASK, n_act >= 1), m2 = __ballot_sync(FULLMASK, n_act >= 2);
      const unsigned t1 = __ballot_sync(FULLMASK, n >= 1), t2 = __ballot_sync(FULLMASK, n >= 2);
      const unsigned lt = (1u << lane) - 1;
      int slot = ncon + __popc(m1 & lt) + __popc(m2 & lt);
      int dslot = ncontot + __popc(t1 & lt) + __popc(t2 & lt);
      for (int i = 0; i < n; i++) {
        if (rc[i].dist < T(0)) {
          if (slot < D.maxcon) contact_rows(slot, rc[i], pa, e, ssign, dbg, dslot + i);
          else flags |= SG_ST_CON_FULL_BIT;
          slot++;
        }
      }
      ncon += __popc(m1) + __popc(m2);
      ncontot += __popc(t1) + __popc(t2);
    }
    if (ncon > D.maxcon) ncon = D.maxcon;
    touch = __reduce_or_sync(FULLMASK, touch);
    flags = __reduce_or_sync(FULLMASK, flags);
    if (lane == 0) { misc(MI_NCON) = ncon; misc(MI_TOUCH) = touch; misc(MI_NCONTOT) = ncontot; misc(MI_STATUS) |= flags; misc(MI_NCAND) = ncand; }
    __syncwarp();
    // ---- Gauss-Seidel levels of the contact blocks: a block waits for the previous block that shares
    // its finger chain or its slider (exactly the sequential order of mj_solPGS) ----
    if (lane == 0) {
      int lastchain[MAXCHAIN];
#pragma unroll
      for (int c = 0; c < MAXCHAIN; c++) lastchain[c] = 0;
      int maxlev = 0;
      for (int i = 0; i < ncon; i++) {
        const int c = smi[L.i_con + 3 * i], e = smi[L.i_con + 3 * i + 1];
        int lv = 0;
        if (c >= 0) lv = lastchain[c];
        if (e >= 0) for (int j = i - 1; j >= 0; j--) if (smi[L.i_con + 3 * j + 1] == e) { const int l2 = smi[L.i_con + 3 * j + 2]; if (l2 > lv) lv = l2; break; }
        lv += 1;
        smi[L.i_con + 3 * i + 2] = lv;
        if (c >= 0) lastchain[c] = lv;
        if (lv > maxlev) maxlev = lv;
      }
      misc(MI_MAXLEV) = maxlev;
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------------------------------
  // equality / tendon / limit rows and smooth dynamics of the shell
  // ------------------------------------------------------------------------------------------
  __device__ void rows_and_smooth() {
    const int nfd = D.nfd, ns = D.ns;
    // volume tendon: L = sum c_e q_e, Ldot = sum c_e v_e
    T Ls = 0, Lv = 0;
    for (int e = lane; e < ns; e += 32) { const T tc = tab(D.o_sl_tc)[e]; Ls += tc * q()[nfd + e]; Lv += tc * v()[nfd + e]; }
    Ls = warp_sum(Ls); Lv = warp_sum(Lv);
    const T Ft = -ten_stiffness() * (Ls - T(D.ten_lspring)) - ten_damping() * Lv;
    // sliders: qfrc_smooth = passive - bias  (bias = -m axis.g for a slider on a static parent)
    for (int e = lane; e < ns; e += 32) {
      const T qe = q()[nfd + e], ve = v()[nfd + e], m = tab(D.o_sl_m)[e];
      const T* ax = tab(D.o_sl_axis) + 3 * e;
      T f = -stiffness(e) * qe - damping(e) * ve;
      f += tab(D.o_sl_tc)[e] * Ft;
      f -= -(m * (ax[0] * T(D.g[0]) + ax[1] * T(D.g[1]) + ax[2] * T(D.g[2])));
      qs()[nfd + e] = f;
    }
    // joint-equality rows in schedule order
    for (int p = lane; p < D.nrow; p += 32) {
      const int d1 = itab(D.io_row_d1)[p], d2 = itab(D.io_row_d2)[p];
      T pos = q()[nfd + d1], vel = v()[nfd + d1], diag = tab(D.o_sl_iw)[d1];
      if (d2 >= 0) { pos -= q()[nfd + d2]; vel -= v()[nfd + d2]; diag += tab(D.o_sl_iw)[d2]; }
      const T imp = impedance<T>(D.eqj_solimp, pos);
      sm[L.rR + p] = tmax(T(SG_MINVAL), (T(1) - imp) * diag / imp);
      sm[L.raref + p] = -T(D.eqj_B) * vel - T(D.eqj_K) * imp * pos;
    }
    if (lane == 0) {
      const T pos = Ls - T(D.ten_l0);
      const T imp = impedance<T>(D.eqt_solimp, pos);
      const T R = tmax(T(SG_MINVAL), (T(1) - imp) * T(D.ten_iw) / imp);
      T* tn = sm + L.ten;
      tn[1] = -T(D.eqt_B) * Lv - T(D.eqt_K) * imp * pos; tn[2] = R; tn[4] = Ls; tn[5] = Lv; tn[6] = Ft;
    }
    // tendon row diagonal: sum c_e^2 / m_e + R
    T As = 0;
    for (int e = lane; e < ns; e += 32) { const T tc = tab(D.o_sl_tc)[e]; As += tc * tc / tab(D.o_sl_m)[e]; }
    As = warp_sum(As);
    __syncwarp();
    if (lane == 0) sm[L.ten + 3] = As + sm[L.ten + 2];
    // joint limits of the finger dofs, lower then upper, in joint order (mj_instantiateLimit)
    bool act = false; T dist = 0, sgn = 0;
    int cchain = 0, jl = 0;
    if (lane < nfd) {
      for (int c = 0; c < D.nchain; c++) if (lane >= D.chain_dof0[c] && lane < D.chain_dof0[c] + D.ncd[c]) { cchain = c; jl = lane - D.chain_dof0[c]; }
      const T* cd = tab(D.o_chain + cchain * CH_STRIDE + CH_DOF + jl * CD_STRIDE);
      if (cd[CD_LIMITED] != T(0)) {
        const T qq = q()[lane];
        const T dlo = qq - cd[CD_LO], dhi = cd[CD_HI] - qq;
        if (dlo < T(0)) { act = true; dist = dlo; sgn = 1; }
        else if (dhi < T(0)) { act = true; dist = dhi; sgn = -1; }
      }
    }
    const unsigned lm = __ballot_sync(FULLMASK, act);
    if (act) {
      const int slot = __popc(lm & ((1u << lane) - 1));
      const T* cd = tab(D.o_chain + cchain * CH_STRIDE + CH_DOF + jl * CD_STRIDE);
      const T imp = impedance<T>(D.lim_solimp, dist);
      const T R = tmax(T(SG_MINVAL), (T(1) - imp) * cd[CD_IW] / imp);
      const T vel = sgn * v()[lane];
      sm[L.lim + 3 * slot + 1] = -T(D.lim_B) * vel - T(D.lim_K) * imp * dist;
      sm[L.lim + 3 * slot + 2] = R;
      smi[L.i_lim + 2 * slot] = lane; smi[L.i_lim + 2 * slot + 1] = sgn > T(0) ? 1 : -1;
    }
    if (lane == 0) misc(MI_NLIM) = __popc(lm);
    __syncwarp();
  }

  __device__ __forceinline__ int chain_of(int dof) const {
    int c = 0;
    for (int k = 1; k < D.nchain; k++) if (dof >= D.chain_dof0[k]) c = k;
    return c;
  }

  // jtf = J^T f for all dofs (deterministic gathers; contact scatter is serial in contact order)
  __device__ void compute_jtf() {
    const int nfd = D.nfd, ns = D.ns;
    T* jtf = sm + L.jtf;
    const T tf = sm[L.ten];
    for (int e = lane; e < ns; e += 32) {
      T s = 0;
      const int* dr = itab(D.io_dof_rows) + e * MAXDOFROWS;
#pragma unroll
      for (int k = 0; k < MAXDOFROWS; k++) {
        const int code = dr[k];
        if (code >= 0) { const T f = sm[L.rf + (code >> 1)]; s += (code & 1) ? -f : f; }
      }
      s += tab(D.o_sl_tc)[e] * tf;
      jtf[nfd + e] = s;
    }
    if (lane < nfd) jtf[lane] = 0;
    __syncwarp();
    if (lane == 0) {
      const int nlim = misc(MI_NLIM), ncon = misc(MI_NCON);
      for (int i = 0; i < nlim; i++) jtf[smi[L.i_lim + 2 * i]] += T(smi[L.i_lim + 2 * i + 1]) * sm[L.lim + 3 * i];
      for (int i = 0; i < ncon; i++) {
        const int c = smi[L.i_con + 3 * i], e = smi[L.i_con + 3 * i + 1];
        const T* f = sm + L.c_f + 3 * i;
        if (e >= 0) { const T* nsv = sm + L.c_ns + 3 * i; jtf[nfd + e] += nsv[0] * f[0] + nsv[1] * f[1] + nsv[2] * f[2]; }
        if (c >= 0) {
          const T* jg = sm + L.c_jg + 12 * i; const int d0 = D.chain_dof0[c];
          for (int jj = 0; jj < D.ncd[c]; jj++) jtf[d0 + jj] += jg[jj] * f[0] + jg[4 + jj] * f[1] + jg[8 + jj] * f[2];
        }
      }
    }
    __syncwarp();
  }

  // a = qacc_smooth + M^-1 jtf (or qacc_smooth alone when !use)
  __device__ void set_qacc(bool use) {
    const int nfd = D.nfd;
    const T* jtf = sm + L.jtf;
    for (int e = lane; e < D.ns; e += 32) {
      const T iw = T(1) / tab(D.o_sl_m)[e];
      a()[nfd + e] = qs()[nfd + e] * iw + (use ? jtf[nfd + e] * iw : T(0));
    }
    if (lane < nfd) {
      const int c = chain_of(lane), jl = lane - D.chain_dof0[c];
      T s = 0;
      if (use) for (int jj = 0; jj < D.ncd[c]; jj++) s += sm[L.g_minv + 16 * c + 4 * jl + jj] * jtf[D.chain_dof0[c] + jj];
      a()[lane] = qs()[lane] + s;
    }
    __syncwarp();
  }

  // warm start: forces from qacc_warmstart (currently in a()), kept only if the dual
  // cost f.b + 0.5 f'AR f is not positive
  __device__ void warmstart() {
    const int nfd = D.nfd, ns = D.ns;
    T cost = 0;
    for (int p = lane; p < D.nrow; p += 32) {
      const int d1 = itab(D.io_row_d1)[p], d2 = itab(D.io_row_d2)[p];
      T ja = a()[nfd + d1], js = qs()[nfd + d1] / tab(D.o_sl_m)[d1];
      if (d2 >= 0) { ja -= a()[nfd + d2]; js -= qs()[nfd + d2] / tab(D.o_sl_m)[d2]; }
      const T ar = sm[L.raref + p], R = sm[L.rR + p];
      const T f = -(T(1) / R) * (ja - ar);
      sm[L.rf + p] = f;
      cost += f * (js - ar) + T(0.5) * R * f * f;
    }
    {  // tendon row
      T ja = 0, js = 0;
      for (int e = lane; e < ns; e += 32) { const T tc = tab(D.o_sl_tc)[e]; ja += tc * a()[nfd + e]; js += tc * qs()[nfd + e] / tab(D.o_sl_m)[e]; }
      ja = warp_sum(ja); js = warp_sum(js);
      if (lane == 0) {
        T* tn = sm + L.ten;
        const T f = -(T(1) / tn[2]) * (ja - tn[1]);
        tn[0] = f;
        cost += f * (js - tn[1]) + T(0.5) * tn[2] * f * f;
      }
    }
    const int nlim = misc(MI_NLIM), ncon = misc(MI_NCON);
    if (lane < nlim) {
      const int dof = smi[L.i_lim + 2 * lane]; const T sgn = T(smi[L.i_lim + 2 * lane + 1]);
      T* lr = sm + L.lim + 3 * lane;
      const T jar = sgn * a()[dof] - lr[1];
      const T f = jar >= T(0) ? T(0) : -(T(1) / lr[2]) * jar;
      lr[0] = f;
      cost += f * (sgn * qs()[dof] - lr[1]) + T(0.5) * lr[2] * f * f;
    }
    for (int i = lane; i < ncon; i += 32) {
      const int c = smi[L.i_con + 3 * i], e = smi[L.i_con + 3 * i + 1];
      const T* jg = sm + L.c_jg + 12 * i; const T* nsv = sm + L.c_ns + 3 * i; const T* ar = sm + L.c_aref + 3 * i;
      const T R0 = sm[L.c_R + 2 * i], R1 = sm[L.c_R + 2 * i + 1];
      T jar[3], b[3];
#pragma unroll
      for (int r = 0; r < 3; r++) {
        T sa = 0, sb = 0;
        if (e >= 0) { sa = nsv[r] * a()[nfd + e]; sb = nsv[r] * (qs()[nfd + e] / tab(D.o_sl_m)[e]); }
        if (c >= 0) for (int jj = 0; jj < D.ncd[c]; jj++) { sa += jg[4 * r + jj] * a()[D.chain_dof0[c] + jj]; sb += jg[4 * r + jj] * qs()[D.chain_dof0[c] + jj]; }
        jar[r] = sa - ar[r]; b[r] = sb - ar[r];
      }
      const T frc = T(D.con_fr), mu = frc * tsqrt(R1 / R0);
      const T Rr[3] = {R0, R1, R1};
      T f[3] = {-(T(1) / R0) * jar[0], -(T(1) / R1) * jar[1], -(T(1) / R1) * jar[2]};
      const T U0 = jar[0] * mu, U1 = jar[1] * frc, U2 = jar[2] * frc;
      const T N = U0, Tn = tsqrt(U1 * U1 + U2 * U2);
      if (N >= mu * Tn || (Tn <= T(0) && N >= T(0))) { f[0] = f[1] = f[2] = 0; }
      else if (mu * N + Tn <= T(0) || (Tn <= T(0) && N < T(0))) { }
      else {
        const T Dm = (T(1) / R0) / (mu * mu * (T(1) + mu * mu)), NT = N - mu * Tn;
        f[0] = -Dm * NT * mu;
        f[1] = -f[0] / Tn * U1 * frc; f[2] = -f[0] / Tn * U2 * frc;
      }
#pragma unroll
      for (int r = 0; r < 3; r++) { sm[L.c_f + 3 * i + r] = f[r]; cost += f[r] * b[r] + T(0.5) * Rr[r] * f[r] * f[r]; }
    }
    __syncwarp();
    compute_jtf();
    // 0.5 f' J M^-1 J' f = 0.5 jtf . (M^-1 jtf)
    const T* jtf = sm + L.jtf;
    for (int e = lane; e < ns; e += 32) { const T x = jtf[nfd + e]; cost += T(0.5) * x * x / tab(D.o_sl_m)[e]; }
    if (lane < nfd) {
      const int c = chain_of(lane), jl = lane - D.chain_dof0[c];
      T s = 0;
      for (int jj = 0; jj < D.ncd[c]; jj++) s += sm[L.g_minv + 16 * c + 4 * jl + jj] * jtf[D.chain_dof0[c] + jj];
      cost += T(0.5) * jtf[lane] * s;
    }
    cost = warp_sum(cost);
    const bool keep = !(cost > T(0));
    if (!keep) {
      for (int p = lane; p < D.nrow; p += 32) sm[L.rf + p] = 0;
      if (lane == 0) sm[L.ten] = 0;
      if (lane < nlim) sm[L.lim + 3 * lane] = 0;
      for (int i = lane; i < 3 * ncon; i += 32) sm[L.c_f + i] = 0;
    }
    __syncwarp();
    set_qacc(keep);
  }

  // one elliptic contact block (mj_solPGS inner body, dim 3)
  __device__ T contact_block(int i) {
    const int nfd = D.nfd;
    const int c = smi[L.i_con + 3 * i], e = smi[L.i_con + 3 * i + 1];
    const T* jg = sm + L.c_jg + 12 * i; const T* nsv = sm + L.c_ns + 3 * i; const T* ar = sm + L.c_aref + 3 * i;
    const T R0 = sm[L.c_R + 2 * i], R1 = sm[L.c_R + 2 * i + 1];
    const T* Ap = sm + L.c_A + 6 * i;
    const T A00 = Ap[0], A01 = Ap[1], A02 = Ap[2], A11 = Ap[3], A12 = Ap[4], A22 = Ap[5];
    T* fp = sm + L.c_f + 3 * i;
    const T old0 = fp[0], old1 = fp[1], old2 = fp[2];
    T ag[MAXCD]; T ae = 0; int d0 = 0, ncd = 0;
    if (c >= 0) { d0 = D.chain_dof0[c]; ncd = D.ncd[c]; }
#pragma unroll
    for (int jj = 0; jj < MAXCD; jj++) ag[jj] = (jj < ncd) ? a()[d0 + jj] : T(0);
    if (e >= 0) ae = a()[nfd + e];
    T res[3];
    const T Rr[3] = {R0, R1, R1};
    const T fo[3] = {old0, old1, old2};
#pragma unroll
    for (int r = 0; r < 3; r++) {
      T s = nsv[r] * ae;
#pragma unroll
      for (int jj = 0; jj < MAXCD; jj++) s += jg[4 * r + jj] * ag[jj];
      res[r] = s - ar[r] + Rr[r] * fo[r];
    }
    T f0 = old0, f1 = old1, f2 = old2;
    if (f0 < T(SG_MINVAL)) {
      f0 -= res[0] / A00;
      if (f0 < T(0)) f0 = 0;
      f1 = 0; f2 = 0;
    } else {
      const T v0 = f0, v1 = f1, v2 = f2;
      const T w0 = A00 * v0 + A01 * v1 + A02 * v2, w1 = A01 * v0 + A11 * v1 + A12 * v2, w2 = A02 * v0 + A12 * v1 + A22 * v2;
      const T denom = v0 * w0 + v1 * w1 + v2 * w2;
      if (denom >= T(SG_MINVAL)) {
        T x = -(v0 * res[0] + v1 * res[1] + v2 * res[2]) / denom;
        if (f0 + x * v0 < T(0)) x = T(-1);
        f0 += x * v0; f1 += x * v1; f2 += x * v2;
      }
    }
    // friction update with the normal force fixed
    {
      T bc[2];
      bc[0] = res[1] - (A11 * old1 + A12 * old2) + A01 * (f0 - old0);
      bc[1] = res[2] - (A12 * old1 + A22 * old2) + A02 * (f0 - old0);
      if (f0 < T(SG_MINVAL)) { f1 = 0; f2 = 0; }
      else {
        const T frc = T(D.con_fr);
        T vv[2];
        const int active = qcqp2<T>(vv, A11, A12, A22, bc, frc, frc, f0);
        if (active) {
          T s = vv[0] * vv[0] / (frc * frc) + vv[1] * vv[1] / (frc * frc);
          s = tsqrt(f0 * f0 / tmax(T(SG_MINVAL), s));
          vv[0] *= s; vv[1] *= s;
        }
        f1 = vv[0]; f2 = vv[1];
      }
    }
    // cost change, revert if positive
    T d0f = f0 - old0, d1f = f1 - old1, d2f = f2 - old2;
    T change = T(0.5) * (d0f * (A00 * d0f + A01 * d1f + A02 * d2f) + d1f * (A01 * d0f + A11 * d1f + A12 * d2f) + d2f * (A02 * d0f + A12 * d1f + A22 * d2f))
             + d0f * res[0] + d1f * res[1] + d2f * res[2];
    if (change > T(1e-10)) { f0 = old0; f1 = old1; f2 = old2; d0f = d1f = d2f = 0; change = 0; }
    fp[0] = f0; fp[1] = f1; fp[2] = f2;
    // qacc += M^-1 J^T delta
    if (d0f != T(0) || d1f != T(0) || d2f != T(0)) {
      if (e >= 0) a()[nfd + e] = ae + (nsv[0] * d0f + nsv[1] * d1f + nsv[2] * d2f) / tab(D.o_sl_m)[e];
      if (c >= 0) {
        T gv[MAXCD];
#pragma unroll
        for (int jj = 0; jj < MAXCD; jj++) gv[jj] = jg[jj] * d0f + jg[4 + jj] * d1f + jg[8 + jj] * d2f;
#pragma unroll
        for (int ii = 0; ii < MAXCD; ii++) {
          if (ii >= ncd) break;
          T s = 0;
#pragma unroll
          for (int jj = 0; jj < MAXCD; jj++) s += sm[L.g_minv + 16 * c + 4 * ii + jj] * gv[jj];
          a()[d0 + ii] = ag[ii] + s;
        }
      }
    }
    return change;
  }

  // projected Gauss-Seidel (mj_solPGS) in MuJoCo's row order
  __device__ void pgs() {
    const int nfd = D.nfd, ns = D.ns;
    const int* lev_start = itab(D.io_lev_start);
    const int* row_d1 = itab(D.io_row_d1);
    const int* row_d2 = itab(D.io_row_d2);
    const int nlim = misc(MI_NLIM), ncon = misc(MI_NCON), maxlev = misc(MI_MAXLEV);
    T* av = a() + nfd;
    int iter = 0;
    while (iter < D.iters) {
      T impr = 0;
      // ---- equality block, level by level ----
      int p0 = lev_start[0];
      for (int lv = 0; lv < D.nlev; lv++) {
        const int p1 = lev_start[lv + 1];
        const int p = p0 + lane;
        if (p < p1) {
          const int d1 = row_d1[p], d2 = row_d2[p];
          const T iw1 = T(1) / tab(D.o_sl_m)[d1];
          T a1 = av[d1], a2 = 0, iw2 = 0;
          if (d2 >= 0) { a2 = av[d2]; iw2 = T(1) / tab(D.o_sl_m)[d2]; }
          const T f = sm[L.rf + p], R = sm[L.rR + p], ar = sm[L.raref + p];
          const T A = iw1 + iw2 + R;
          const T res = (a1 - a2) - ar + R * f;
          T fn = f - res / A;
          T dl = fn - f;
          T change = T(0.5) * dl * dl * A + dl * res;
          if (change > T(1e-10)) { fn = f; dl = 0; change = 0; }
          impr -= change;
          sm[L.rf + p] = fn;
          av[d1] = a1 + iw1 * dl;
          if (d2 >= 0) av[d2] = a2 - iw2 * dl;
        }
        p0 = p1;
        __syncwarp();
      }
      // ---- volume-tendon row: dense over the shell, warp-shuffle reduction ----
      {
        T s = 0;
        for (int e = lane; e < ns; e += 32) s += tab(D.o_sl_tc)[e] * av[e];
        s = warp_sum(s);
        const T* tn = sm + L.ten;
        const T f = tn[0], ar = tn[1], R = tn[2], A = tn[3];
        const T res = s - ar + R * f;
        T fn = f - res / A;
        T dl = fn - f;
        T change = T(0.5) * dl * dl * A + dl * res;
        if (change > T(1e-10)) { fn = f; dl = 0; change = 0; }
        if (lane == 0) impr -= change;
        __syncwarp();
        if (lane == 0) sm[L.ten] = fn;
        if (dl != T(0)) for (int e = lane; e < ns; e += 32) av[e] += tab(D.o_sl_tc)[e] / tab(D.o_sl_m)[e] * dl;
        __syncwarp();
      }
      // ---- joint limits (finger dofs), sequential ----
      if (lane == 0) {
        for (int i = 0; i < nlim; i++) {
          const int dof = smi[L.i_lim + 2 * i]; const T sgn = T(smi[L.i_lim + 2 * i + 1]);
          const int c = chain_of(dof), jl = dof - D.chain_dof0[c];
          T* lr = sm + L.lim + 3 * i;
          const T f = lr[0], ar = lr[1], R = lr[2];
          const T A = sm[L.g_minv + 16 * c + 4 * jl + jl] + R;
          const T res = sgn * a()[dof] - ar + R * f;
          T fn = f - res / A;
          if (fn < T(0)) fn = 0;
          T dl = fn - f;
          T change = T(0.5) * dl * dl * A + dl * res;
          if (change > T(1e-10)) { fn = f; dl = 0; change = 0; }
          impr -= change;
          lr[0] = fn;
          if (dl != T(0)) for (int ii = 0; ii < D.ncd[c]; ii++) a()[D.chain_dof0[c] + ii] += sm[L.g_minv + 16 * c + 4 * ii + jl] * sgn * dl;
        }
      }
      __syncwarp();
      // ---- elliptic contact blocks by dependency level ----
      for (int lv = 1; lv <= maxlev; lv++) {
        for (int i = lane; i < ncon; i += 32)
          if (smi[L.i_con + 3 * i + 2] == lv) impr -= contact_block(i);
        __syncwarp();
      }
      impr = warp_sum(impr) * T(D.impr_scale);
      iter++;
      if (impr < T(D.tol)) break;
    }
    if (lane == 0) misc(MI_ITERS) = iter;
  }

  // ------------------------------------------------------------------------------------------
  // mj_forward for this world; returns true if qacc is bad
  // ------------------------------------------------------------------------------------------
  __device__ bool forward() {
    if (lane < D.nchain) gripper(lane);
    __syncwarp();
    collide();
    rows_and_smooth();
    warmstart();
    pgs();
    __syncwarp();
    compute_jtf();
    set_qacc(true);
    // accelerometers (mj_sensorAcc): R_site^T (Jv qacc + bias - g)
    if (lane < D.nsens) {
      const T* se = tab(D.o_sens + lane * SE_STRIDE);
      if ((int)se[SE_TYPE] == SENS_ACCEL) {
        const int c = (int)se[SE_CHAIN], d0 = D.chain_dof0[c], adr = (int)se[SE_ADR];
        T acc[3] = {sm[L.s_ab + 3 * lane], sm[L.s_ab + 3 * lane + 1], sm[L.s_ab + 3 * lane + 2]};
        for (int jj = 0; jj < D.ncd[c]; jj++) {
          const T aj = a()[d0 + jj];
#pragma unroll
          for (int k = 0; k < 3; k++) acc[k] += sm[L.s_jv + 12 * lane + 3 * jj + k] * aj;
        }
        T o[3]; matTvec3(o, sm + L.s_rot + 9 * lane, acc);
#pragma unroll
        for (int k = 0; k < 3; k++) sm[L.sens + adr + k] = o[k];
      }
    }
    bool badacc = false;
    for (int i = lane; i < D.nv; i += 32) if (!(tabs(a()[i]) <= T(SG_MAXVAL))) badacc = true;
    __syncwarp();
    return __any_sync(FULLMASK, badacc);
  }

  // mj_Euler with implicit joint damping: (M + h diag(d)) qacc' = qfrc_smooth + J^T f
  __device__ void euler() {
    const int nfd = D.nfd; const T h = T(D.h);
    const T* jtf = sm + L.jtf;
    for (int e = lane; e < D.ns; e += 32) {
      const T m = tab(D.o_sl_m)[e];
      const T qa = (qs()[nfd + e] + jtf[nfd + e]) / (m + h * damping(e));
      const T vn = v()[nfd + e] + h * qa;
      v()[nfd + e] = vn; q()[nfd + e] += h * vn;
    }
    if (lane < nfd) { const T vn = v()[lane] + h * a()[lane]; v()[lane] = vn; q()[lane] += h * vn; }
    if (lane < D.nu) K.act[(size_t)w * D.nu + lane] += h * sm[L.g_actdot + lane];
    __syncwarp();
  }

  // mj_step
  __device__ void step() {
    bool badpv = false;
    for (int i = lane; i < D.nv; i += 32) if (!(tabs(q()[i]) <= T(SG_MAXVAL)) || !(tabs(v()[i]) <= T(SG_MAXVAL))) badpv = true;
    if (__any_sync(FULLMASK, badpv)) { if (lane == 0) misc(MI_STATUS) |= 1; reset_state(); }
    for (int pass = 0; pass < 2; pass++) {
      const bool bad = forward();
      if (!bad) break;
      if (lane == 0) misc(MI_STATUS) |= 1;
      reset_state();
    }
    euler();
  }

  __device__ void debug_dump() {
    // header: [0]=ncon_total [1]=nefc [2]=iters [3]=ncon(rows) [4]=nlim [5]=maxlev [6]=ncand ; contacts at 64+16*i ;
    // then at 64+16*maxcon*2: qacc[nv], qacc_smooth-ish skipped, efc_force / aref / R in MuJoCo row order
    double* o = K.debug_out;
    if (!o) return;
    const int ncon = misc(MI_NCON), nlim = misc(MI_NLIM);
    const int nefc = D.nrow + 1 + nlim + 3 * ncon;
    if (lane == 0) { o[0] = misc(MI_NCONTOT); o[1] = nefc; o[2] = misc(MI_ITERS); o[3] = ncon; o[4] = nlim; o[5] = misc(MI_MAXLEV); o[6] = misc(MI_NCAND); }
    const int base = 64 + 16 * 2 * D.maxcon;
    for (int i = lane; i < D.nv; i += 32) if (base + i < K.debug_cap) o[base + i] = (double)a()[i];
    const int eb = base + D.nv;
    // schedule position -> MuJoCo row id needs the host's table; rows are dumped in schedule order here
    for (int p = lane; p < D.nrow; p += 32) {
      if (eb + 3 * nefc > K.debug_cap) break;
      o[eb + p] = (double)sm[L.rf + p]; o[eb + nefc + p] = (double)sm[L.raref + p]; o[eb + 2 * nefc + p] = (double)sm[L.rR + p];
    }
    if (lane == 0 && eb + 3 * nefc <= K.debug_cap) {
      int r = D.nrow;
      o[eb + r] = (double)sm[L.ten]; o[eb + nefc + r] = (double)sm[L.ten + 1]; o[eb + 2 * nefc + r] = (double)sm[L.ten + 2]; r++;
      for (int i = 0; i < nlim; i++, r++) { o[eb + r] = (double)sm[L.lim + 3 * i]; o[eb + nefc + r] = (double)sm[L.lim + 3 * i + 1]; o[eb + 2 * nefc + r] = (double)sm[L.lim + 3 * i + 2]; }
      for (int i = 0; i < ncon; i++)
        for (int k = 0; k < 3; k++, r++) {
          o[eb + r] = (double)sm[L.c_f + 3 * i + k]; o[eb + nefc + r] = (double)sm[L.c_aref + 3 * i + k];
          o[eb + 2 * nefc + r] = (double)sm[L.c_R + 2 * i + (k ? 1 : 0)];
        }
    }
  }
};

// ---------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(32) sg_step_kernel(const __grid_constant__ KArgs<T> K) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  T* sm = reinterpret_cast<T*>(smem_raw);
  const PlanDims& D = K.D;
  const int lane = threadIdx.x;
  for (int w = blockIdx.x; w < K.nworlds; w += gridDim.x) {
    World<T> W(K, sm, w);
    W.load_params();
    if (lane == 0) W.misc(MI_STATUS) = 0;
    const size_t sb = (size_t)w * D.nv;
    if (!K.rollout) {
      for (int i = lane; i < D.nv; i += 32) { W.q()[i] = K.qpos[sb + i]; W.v()[i] = K.qvel[sb + i]; W.a()[i] = K.warm[sb + i]; }
      __syncwarp();
      if (K.integrate) { for (int s = 0; s < K.nsub; s++) W.step(); }
      else {
        // mj_forward: sensors/contacts refreshed, nothing integrated, warm start left untouched
        W.forward();
        if (w == K.debug_world) W.debug_dump();
        __syncwarp();
        for (int i = lane; i < D.nv; i += 32) W.a()[i] = K.warm[sb + i];
        __syncwarp();
      }
      if (K.integrate && w == K.debug_world) W.debug_dump();
      for (int i = lane; i < D.nv; i += 32) { K.qpos[sb + i] = W.q()[i]; K.qvel[sb + i] = W.v()[i]; K.warm[sb + i] = W.a()[i]; }
      if (K.sens_out && lane < D.nsd) K.sens_out[(size_t)w * D.nsd + lane] = sm[K.L.sens + lane];
      if (K.touch_out && lane == 0) K.touch_out[w] = W.misc(MI_TOUCH);
    } else {
      // whole episode on-chip (create_dataset.log_into_file, ref: create_dataset.py:33-60)
      W.reset_state();
      for (int s = 0; s < K.sim_start; s++) W.step();
      for (int t = 0; t < K.nrows; t++) {
        if (K.ctrl_event[t] && lane < D.nu) K.ctrl[(size_t)w * D.nu + lane] = T(K.ctrl_value[t * D.nu + lane]);
        __syncwarp();
        for (int s = 0; s < K.sim_step; s++) W.step();
        if (lane < D.nsd) K.sens_out[((size_t)w * K.nrows + t) * D.nsd + lane] = sm[K.L.sens + lane];
        if (K.touch_out && lane == 0) K.touch_out[(size_t)w * K.nrows + t] = W.misc(MI_TOUCH);
      }
      for (int i = lane; i < D.nv; i += 32) { K.qpos[sb + i] = W.q()[i]; K.qvel[sb + i] = W.v()[i]; K.warm[sb + i] = W.a()[i]; }
    }
    __syncwarp();
    if (lane == 0 && W.misc(MI_STATUS)) atomicOr(&K.status[w], W.misc(MI_STATUS));
    __syncwarp();
  }
}

}  // namespace sg
