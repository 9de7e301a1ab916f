/* sg_oracle.c -- TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or called from the
 * product path (soft-grip_b200/); only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may use it, and only as the checker / CPU baseline.
 *
 * What it is: an fp64, single-world, single-thread CPU restatement of the physics step the
 * reference executes through mujoco-py on its data/gripper models:
 *     ManEnv.step  -> 7 x sim.step()  == mj_step          (ref: environment/manenv.py:44-53)
 *     ManEnv.reset -> sim.reset(); sim.forward(); step()  (ref: environment/manenv.py:55-63)
 *     sensor read-out sim.data.sensordata / contacts      (ref: environment/manenv.py:65-85)
 *     episode protocol of create_dataset.log_into_file     (ref: create_dataset.py:33-60)
 *
 * The arithmetic of that path lives in the third-party MuJoCo 2.x C engine (un-vendored, no version
 * pinned by the reference; mujoco-py 2.0.x / MuJoCo 2.0-2.1 by date), which is absent from
 * /root/reference and not installable here.  This file therefore restates MuJoCo's *published*
 * algorithm for exactly the feature subset of the reference's MJCF files, stage by stage, following
 * SURVEY.md Appendix A (each function names the MuJoCo stage it restates).
 *
 * PARITY UNPINNED: the reference holds no tests, golden vectors or stored trajectories for this
 * path and the real engine cannot be run here, so this oracle is pinned only by analytic known-
 * answer tests, the independently derived compile-time constants of SURVEY App. D, physical
 * invariants, and a dense "literal efc_AR" PGS mode that cross-checks the matrix-free solver.
 * Two pieces are *defined here* rather than recalled: (i) the capsule-box narrowphase (MuJoCo's
 * mjc_CapsuleBox case analysis is replaced by: closest point of the segment to the box (root of the
 * monotone derivative of the squared distance; if the segment enters the box, its inside end),
 * sphere-box there, plus a second sphere-box at the far end of the segment when that end is also
 * within margin), and (ii) the PGS keeps qacc = qacc_smooth +
 * M^-1 J^T f incrementally ("matrix-free"), algebraically identical to res = b + AR f.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off).
 */
#include "sg_oracle.h"
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define MINVAL 1e-15
#define MAXVAL 1e10
#define MINIMP 0.0001
#define MAXIMP 0.9999

enum { GEOM_PLANE = 0, GEOM_SPHERE = 2, GEOM_CAPSULE = 3, GEOM_BOX = 6 };
enum { JNT_SLIDE = 2, JNT_HINGE = 3 };
enum { EQ_JOINT = 2, EQ_TENDON = 3 };
enum { TEN_FIXED = 0, TEN_SPATIAL = 1 };
enum { SENS_ACCEL = 0, SENS_GYRO = 1 };
enum { CNSTR_EQUALITY = 0, CNSTR_LIMIT_JOINT = 3, CNSTR_CONTACT_ELLIPTIC = 7 };

/* ------------------------------------------------------------------------------------------ */
/* model                                                                                      */
/* ------------------------------------------------------------------------------------------ */
struct sgo_model {
  void* blob;
  int nbody, njnt, nv, ngeom, nsite, ntendon, nwrap, neq, nu, nsensor, nsensordata, nM;
  double timestep, gravity[3], tolerance, impratio, meaninertia;
  int iterations, nconmax, njmax;
  const int *body_parentid, *body_jntadr, *body_jntnum, *body_dofadr, *body_dofnum, *body_geomadr,
      *body_geomnum, *body_weldid, *body_rootid;
  const double *body_pos, *body_quat, *body_ipos, *body_iquat, *body_mass, *body_inertia,
      *body_invweight0, *body_subtreemass;
  const int *jnt_type, *jnt_bodyid, *jnt_limited;
  const double *jnt_pos, *jnt_axis, *jnt_range, *jnt_stiffness, *jnt_margin, *jnt_solref, *jnt_solimp,
      *qpos0, *qpos_spring;
  const int *dof_bodyid, *dof_parentid, *dof_Madr;
  const double *dof_damping, *dof_invweight0;
  const int *geom_type, *geom_bodyid, *geom_contype, *geom_conaffinity, *geom_condim;
  const double *geom_pos, *geom_quat, *geom_size, *geom_friction, *geom_solref, *geom_solimp,
      *geom_margin, *geom_gap, *geom_solmix, *geom_rbound;
  const int* site_bodyid;
  const double *site_pos, *site_quat;
  const int *tendon_type, *tendon_adr, *tendon_num, *wrap_objid;
  const double *tendon_stiffness, *tendon_damping, *tendon_length0, *tendon_lengthspring,
      *tendon_invweight0, *wrap_prm;
  const int *eq_type, *eq_obj1id, *eq_obj2id;
  const double *eq_data, *eq_solref, *eq_solimp;
  const int* actuator_trnid;
  const double *actuator_gear, *actuator_timeconst, *actuator_gain, *actuator_bias;
  const int *sensor_type, *sensor_objid, *sensor_adr;
  /* derived */
  int* dof_simple;                 /* 1: M row is diagonal only                         */
  int* dof_treeid;                 /* connected component of the dof forest             */
  int ntree, *tree_adr, *tree_num, *tree_dofs, *tree_minvadr, ntreeminv;
  int npair, *pair_g1, *pair_g2;   /* statically filtered geom pairs, MuJoCo contact order */
};

typedef struct { const char* name; int dtype; unsigned count; const void* data; } section;

static int find_section(const void* blob, size_t n, const char* name, section* out) {
  const unsigned char* p = (const unsigned char*)blob;
  if (n < 12 || memcmp(p, "SGM1", 4) != 0) return 0;
  unsigned nsec; memcpy(&nsec, p + 8, 4);
  const unsigned char* e = p + 12;
  for (unsigned i = 0; i < nsec; i++, e += 48) {
    if (strncmp((const char*)e, name, 32) == 0) {
      unsigned dt, cnt; unsigned long long off;
      memcpy(&dt, e + 32, 4); memcpy(&cnt, e + 36, 4); memcpy(&off, e + 40, 8);
      out->name = name; out->dtype = (int)dt; out->count = cnt; out->data = p + off;
      return 1;
    }
  }
  return 0;
}

static const double* secd(const sgo_model* m, size_t n, const char* name, int* cnt, int* ok) {
  section s;
  if (!find_section(m->blob, n, name, &s) || s.dtype != 0) { *ok = 0; if (cnt) *cnt = 0; return NULL; }
  if (cnt) *cnt = (int)s.count;
  return (const double*)s.data;
}
static const int* seci(const sgo_model* m, size_t n, const char* name, int* cnt, int* ok) {
  section s;
  if (!find_section(m->blob, n, name, &s) || s.dtype != 1) { *ok = 0; if (cnt) *cnt = 0; return NULL; }
  if (cnt) *cnt = (int)s.count;
  return (const int*)s.data;
}

static void build_pairs(sgo_model* m);

sgo_model* sgo_model_load(const void* blob, size_t nbytes) {
  sgo_model* m = (sgo_model*)calloc(1, sizeof(sgo_model));
  m->blob = malloc(nbytes);
  memcpy(m->blob, blob, nbytes);
  int ok = 1, c;
  const double* opt = secd(m, nbytes, "opt", &c, &ok);
  if (!ok || c < 12) { sgo_model_free(m); return NULL; }
  m->timestep = opt[0]; m->gravity[0] = opt[1]; m->gravity[1] = opt[2]; m->gravity[2] = opt[3];
  m->iterations = (int)opt[4]; m->tolerance = opt[5]; m->impratio = opt[6]; m->meaninertia = opt[7];
  m->nconmax = (int)opt[8]; m->njmax = (int)opt[9]; m->nM = (int)opt[10];
#define D(f) m->f = secd(m, nbytes, #f, NULL, &ok)
#define I(f) m->f = seci(m, nbytes, #f, NULL, &ok)
  m->body_parentid = seci(m, nbytes, "body_parentid", &m->nbody, &ok);
  I(body_jntadr); I(body_jntnum); I(body_dofadr); I(body_dofnum); I(body_geomadr); I(body_geomnum);
  I(body_weldid); I(body_rootid);
  D(body_pos); D(body_quat); D(body_ipos); D(body_iquat); D(body_mass); D(body_inertia);
  D(body_invweight0); D(body_subtreemass);
  m->jnt_type = seci(m, nbytes, "jnt_type", &m->njnt, &ok);
  I(jnt_bodyid); I(jnt_limited);
  D(jnt_pos); D(jnt_axis); D(jnt_range); D(jnt_stiffness); D(jnt_margin); D(jnt_solref); D(jnt_solimp);
  D(qpos0); D(qpos_spring);
  m->dof_bodyid = seci(m, nbytes, "dof_bodyid", &m->nv, &ok);
  I(dof_parentid); I(dof_Madr); D(dof_damping); D(dof_invweight0);
  m->geom_type = seci(m, nbytes, "geom_type", &m->ngeom, &ok);
  I(geom_bodyid); I(geom_contype); I(geom_conaffinity); I(geom_condim);
  D(geom_pos); D(geom_quat); D(geom_size); D(geom_friction); D(geom_solref); D(geom_solimp);
  D(geom_margin); D(geom_gap); D(geom_solmix); D(geom_rbound);
  m->site_bodyid = seci(m, nbytes, "site_bodyid", &m->nsite, &ok);
  D(site_pos); D(site_quat);
  m->tendon_type = seci(m, nbytes, "tendon_type", &m->ntendon, &ok);
  I(tendon_adr); I(tendon_num);
  m->wrap_objid = seci(m, nbytes, "wrap_objid", &m->nwrap, &ok);
  D(tendon_stiffness); D(tendon_damping); D(tendon_length0); D(tendon_lengthspring);
  D(tendon_invweight0); D(wrap_prm);
  m->eq_type = seci(m, nbytes, "eq_type", &m->neq, &ok);
  I(eq_obj1id); I(eq_obj2id); D(eq_data); D(eq_solref); D(eq_solimp);
  m->actuator_trnid = seci(m, nbytes, "actuator_trnid", &m->nu, &ok);
  D(actuator_gear); D(actuator_timeconst); D(actuator_gain); D(actuator_bias);
  m->sensor_type = seci(m, nbytes, "sensor_type", &m->nsensor, &ok);
  I(sensor_objid); I(sensor_adr);
#undef D
#undef I
  if (!ok || m->njnt != m->nv) { sgo_model_free(m); return NULL; }
  m->nsensordata = 3 * m->nsensor;
  /* dof forest: simple flags, trees */
  int nv = m->nv;
  m->dof_simple = (int*)malloc(sizeof(int) * nv);
  m->dof_treeid = (int*)malloc(sizeof(int) * nv);
  for (int i = 0; i < nv; i++) m->dof_simple[i] = (m->dof_parentid[i] < 0);
  for (int i = 0; i < nv; i++) if (m->dof_parentid[i] >= 0) m->dof_simple[m->dof_parentid[i]] = 0;
  m->ntree = 0;
  for (int i = 0; i < nv; i++) {
    if (m->dof_parentid[i] < 0) m->dof_treeid[i] = m->ntree++;
    else m->dof_treeid[i] = m->dof_treeid[m->dof_parentid[i]];
  }
  m->tree_adr = (int*)calloc(m->ntree + 1, sizeof(int));
  m->tree_num = (int*)calloc(m->ntree, sizeof(int));
  m->tree_dofs = (int*)malloc(sizeof(int) * nv);
  m->tree_minvadr = (int*)calloc(m->ntree, sizeof(int));
  for (int i = 0; i < nv; i++) m->tree_num[m->dof_treeid[i]]++;
  for (int t = 0; t < m->ntree; t++) m->tree_adr[t + 1] = m->tree_adr[t] + m->tree_num[t];
  int* fill = (int*)calloc(m->ntree, sizeof(int));
  for (int i = 0; i < nv; i++) { int t = m->dof_treeid[i]; m->tree_dofs[m->tree_adr[t] + fill[t]++] = i; }
  free(fill);
  m->ntreeminv = 0;
  for (int t = 0; t < m->ntree; t++) { m->tree_minvadr[t] = m->ntreeminv; m->ntreeminv += m->tree_num[t] * m->tree_num[t]; }
  build_pairs(m);
  return m;
}

void sgo_model_free(sgo_model* m) {
  if (!m) return;
  free(m->blob); free(m->dof_simple); free(m->dof_treeid); free(m->tree_adr); free(m->tree_num);
  free(m->tree_dofs); free(m->tree_minvadr); free(m->pair_g1); free(m->pair_g2);
  free(m);
}

int sgo_model_int(const sgo_model* m, const char* k) {
  if (!strcmp(k, "nv")) return m->nv;
  if (!strcmp(k, "nbody")) return m->nbody;
  if (!strcmp(k, "ngeom")) return m->ngeom;
  if (!strcmp(k, "neq")) return m->neq;
  if (!strcmp(k, "nu")) return m->nu;
  if (!strcmp(k, "nsensordata")) return m->nsensordata;
  if (!strcmp(k, "ntendon")) return m->ntendon;
  if (!strcmp(k, "njnt")) return m->njnt;
  if (!strcmp(k, "npair")) return m->npair;
  if (!strcmp(k, "nsite")) return m->nsite;
  if (!strcmp(k, "nM")) return m->nM;
  return -1;
}

/* candidate geom pairs that survive the static filters of mj_collision (SURVEY App. A1 "Collision"):
 * same weld body, parent-child (only when neither weld body is the world), contype/conaffinity.
 * Emitted in MuJoCo's contact order: body pairs (b1<b2) ascending, geoms in body order; inside a
 * pair the geom with the lower type id comes first. */
static void build_pairs(sgo_model* m) {
  int cap = 1024, n = 0;
  int* g1s = (int*)malloc(sizeof(int) * cap);
  int* g2s = (int*)malloc(sizeof(int) * cap);
  for (int b1 = 0; b1 < m->nbody; b1++) {
    if (!m->body_geomnum[b1]) continue;
    for (int b2 = b1 + 1; b2 < m->nbody; b2++) {
      if (!m->body_geomnum[b2]) continue;
      int w1 = m->body_weldid[b1], w2 = m->body_weldid[b2];
      if (w1 == w2) continue;
      int wp1 = m->body_weldid[m->body_parentid[w1]], wp2 = m->body_weldid[m->body_parentid[w2]];
      if (w1 != 0 && w2 != 0 && (w1 == wp2 || w2 == wp1)) continue;
      for (int a = m->body_geomadr[b1]; a < m->body_geomadr[b1] + m->body_geomnum[b1]; a++)
        for (int b = m->body_geomadr[b2]; b < m->body_geomadr[b2] + m->body_geomnum[b2]; b++) {
          if (!((m->geom_contype[a] & m->geom_conaffinity[b]) || (m->geom_contype[b] & m->geom_conaffinity[a]))) continue;
          if (n == cap) { cap *= 2; g1s = (int*)realloc(g1s, sizeof(int) * cap); g2s = (int*)realloc(g2s, sizeof(int) * cap); }
          if (m->geom_type[a] <= m->geom_type[b]) { g1s[n] = a; g2s[n] = b; } else { g1s[n] = b; g2s[n] = a; }
          n++;
        }
    }
  }
  m->npair = n; m->pair_g1 = g1s; m->pair_g2 = g2s;
}

/* ------------------------------------------------------------------------------------------ */
/* world (the mjData of one environment plus its per-world model parameters)                   */
/* ------------------------------------------------------------------------------------------ */
struct sgo_world {
  const sgo_model* m;
  double *jnt_stiffness, *tendon_stiffness, *dof_damping, *tendon_damping, *body_pos;
  int* geom_mask;
  double *qpos, *qvel, *act, *ctrl, *qacc, *qacc_warmstart, time;
  double *xpos, *xquat, *xmat, *xipos, *ximat, *xanchor, *xaxis, *geom_xpos, *geom_xmat, *site_xpos, *site_xmat;
  double *subtree_com, *cinert, *crb, *cdof, *cdof_dot, *cvel, *cacc;
  double *ten_length, *ten_velocity, *ten_J;
  double *qM, *qLD, *qLDiagInv, *tree_Minv;
  double *qfrc_passive, *qfrc_bias, *qfrc_actuator, *qfrc_smooth, *qacc_smooth, *qfrc_constraint;
  double *act_dot, *actuator_force;
  int ncon;
  double *con_dist, *con_pos, *con_frame, *con_friction, *con_solref, *con_solimp, *con_mu, *con_includemargin;
  int *con_geom1, *con_geom2, *con_dim, *con_exclude, *con_efc;
  int nefc, ne, nl, nnz, nnzcap;
  int *efc_type, *efc_id, *efc_rowadr, *efc_rownnz, *efc_colind;
  double *efc_pos, *efc_margin, *efc_diagApprox, *efc_R, *efc_D, *efc_K, *efc_Bd, *efc_imp, *efc_vel,
      *efc_aref, *efc_b, *efc_force, *efc_J, *efc_B;
  double* sensordata;
  double* scratch;   /* 8*nv */
  double* cfrc;      /* 6*nbody */
  double* efc_A;     /* 3 per row: diagonal block of AR */
  int* mark;         /* nv   */
  int status, step_status, touch_mask, solver_iter, dense, cb_single, implicit_tendon;
  double flops;            /* PGS op count of the last forward */
  double flops_total, flops_pgs_total; long steps_total;   /* accumulated over every forward since the last sgo_flops_reset */
  int ncand;               /* geom pairs that reached the narrowphase in the last forward */
};

static double* dalloc(size_t n) { return (double*)calloc(n ? n : 1, sizeof(double)); }
static int* ialloc(size_t n) { return (int*)calloc(n ? n : 1, sizeof(int)); }

sgo_world* sgo_world_create(const sgo_model* m) {
  sgo_world* d = (sgo_world*)calloc(1, sizeof(sgo_world));
  d->m = m;
  int nv = m->nv, nb = m->nbody, nj = m->njnt, ng = m->ngeom, ns = m->nsite, nt = m->ntendon;
  d->jnt_stiffness = dalloc(nj); memcpy(d->jnt_stiffness, m->jnt_stiffness, sizeof(double) * nj);
  d->tendon_stiffness = dalloc(nt); memcpy(d->tendon_stiffness, m->tendon_stiffness, sizeof(double) * nt);
  d->dof_damping = dalloc(nv); memcpy(d->dof_damping, m->dof_damping, sizeof(double) * nv);
  d->tendon_damping = dalloc(nt); memcpy(d->tendon_damping, m->tendon_damping, sizeof(double) * nt);
  d->body_pos = dalloc(3 * nb); memcpy(d->body_pos, m->body_pos, sizeof(double) * 3 * nb);
  d->geom_mask = ialloc(ng);
  d->qpos = dalloc(nv); d->qvel = dalloc(nv); d->act = dalloc(m->nu); d->ctrl = dalloc(m->nu);
  d->qacc = dalloc(nv); d->qacc_warmstart = dalloc(nv);
  d->xpos = dalloc(3 * nb); d->xquat = dalloc(4 * nb); d->xmat = dalloc(9 * nb); d->xipos = dalloc(3 * nb);
  d->ximat = dalloc(9 * nb); d->xanchor = dalloc(3 * nj); d->xaxis = dalloc(3 * nj);
  d->geom_xpos = dalloc(3 * ng); d->geom_xmat = dalloc(9 * ng); d->site_xpos = dalloc(3 * ns); d->site_xmat = dalloc(9 * ns);
  d->subtree_com = dalloc(3 * nb); d->cinert = dalloc(10 * nb); d->crb = dalloc(10 * nb);
  d->cdof = dalloc(6 * nv); d->cdof_dot = dalloc(6 * nv); d->cvel = dalloc(6 * nb); d->cacc = dalloc(6 * nb);
  d->ten_length = dalloc(nt); d->ten_velocity = dalloc(nt); d->ten_J = dalloc((size_t)nt * nv);
  d->qM = dalloc(m->nM); d->qLD = dalloc(m->nM); d->qLDiagInv = dalloc(nv); d->tree_Minv = dalloc(m->ntreeminv);
  d->qfrc_passive = dalloc(nv); d->qfrc_bias = dalloc(nv); d->qfrc_actuator = dalloc(nv);
  d->qfrc_smooth = dalloc(nv); d->qacc_smooth = dalloc(nv); d->qfrc_constraint = dalloc(nv);
  d->act_dot = dalloc(m->nu); d->actuator_force = dalloc(m->nu);
  int nc = m->nconmax;
  d->con_dist = dalloc(nc); d->con_pos = dalloc(3 * nc); d->con_frame = dalloc(9 * nc);
  d->con_friction = dalloc(5 * nc); d->con_solref = dalloc(2 * nc); d->con_solimp = dalloc(5 * nc);
  d->con_mu = dalloc(nc); d->con_includemargin = dalloc(nc);
  d->con_geom1 = ialloc(nc); d->con_geom2 = ialloc(nc); d->con_dim = ialloc(nc); d->con_exclude = ialloc(nc);
  d->con_efc = ialloc(nc);
  int ne = m->njmax;
  d->efc_type = ialloc(ne); d->efc_id = ialloc(ne); d->efc_rowadr = ialloc(ne); d->efc_rownnz = ialloc(ne);
  d->efc_pos = dalloc(ne); d->efc_margin = dalloc(ne); d->efc_diagApprox = dalloc(ne); d->efc_R = dalloc(ne);
  d->efc_D = dalloc(ne); d->efc_K = dalloc(ne); d->efc_Bd = dalloc(ne); d->efc_imp = dalloc(ne);
  d->efc_vel = dalloc(ne); d->efc_aref = dalloc(ne); d->efc_b = dalloc(ne); d->efc_force = dalloc(ne);
  d->nnzcap = m->neq * 2 + nt * nv + 16 * nj + nc * 3 * 24 + 64;
  d->efc_colind = ialloc(d->nnzcap); d->efc_J = dalloc(d->nnzcap); d->efc_B = dalloc(d->nnzcap);
  d->sensordata = dalloc(m->nsensordata);
  d->scratch = dalloc(8 * (size_t)nv + 64); d->mark = ialloc(nv);
  d->cfrc = dalloc(6 * (size_t)nb); d->efc_A = dalloc(3 * (size_t)ne);
  sgo_reset(d);
  return d;
}

void sgo_world_free(sgo_world* d) {
  if (!d) return;
  void* ptrs[] = {d->jnt_stiffness, d->tendon_stiffness, d->dof_damping, d->tendon_damping, d->body_pos, d->geom_mask,
    d->qpos, d->qvel, d->act, d->ctrl, d->qacc, d->qacc_warmstart, d->xpos, d->xquat, d->xmat, d->xipos, d->ximat,
    d->xanchor, d->xaxis, d->geom_xpos, d->geom_xmat, d->site_xpos, d->site_xmat, d->subtree_com, d->cinert, d->crb,
    d->cdof, d->cdof_dot, d->cvel, d->cacc, d->ten_length, d->ten_velocity, d->ten_J, d->qM, d->qLD, d->qLDiagInv,
    d->tree_Minv, d->qfrc_passive, d->qfrc_bias, d->qfrc_actuator, d->qfrc_smooth, d->qacc_smooth, d->qfrc_constraint,
    d->act_dot, d->actuator_force, d->con_dist, d->con_pos, d->con_frame, d->con_friction, d->con_solref, d->con_solimp,
    d->con_mu, d->con_includemargin, d->con_geom1, d->con_geom2, d->con_dim, d->con_exclude, d->con_efc, d->efc_type,
    d->efc_id, d->efc_rowadr, d->efc_rownnz, d->efc_pos, d->efc_margin, d->efc_diagApprox, d->efc_R, d->efc_D, d->efc_K,
    d->efc_Bd, d->efc_imp, d->efc_vel, d->efc_aref, d->efc_b, d->efc_force, d->efc_colind, d->efc_J, d->efc_B,
    d->sensordata, d->scratch, d->mark, d->cfrc, d->efc_A};
  for (size_t i = 0; i < sizeof(ptrs) / sizeof(ptrs[0]); i++) free(ptrs[i]);
  free(d);
}

void sgo_set_jnt_stiffness(sgo_world* d, int j, double k) { if (j >= 0 && j < d->m->njnt) d->jnt_stiffness[j] = k; }
void sgo_set_tendon_stiffness(sgo_world* d, int t, double k) { if (t >= 0 && t < d->m->ntendon) d->tendon_stiffness[t] = k; }
void sgo_set_dof_damping(sgo_world* d, int i, double v) { if (i >= 0 && i < d->m->nv) d->dof_damping[i] = v; }
void sgo_set_tendon_damping(sgo_world* d, int t, double v) { if (t >= 0 && t < d->m->ntendon) d->tendon_damping[t] = v; }
void sgo_set_body_pos(sgo_world* d, int b, const double* p) { if (b > 0 && b < d->m->nbody) memcpy(d->body_pos + 3 * b, p, 3 * sizeof(double)); }
void sgo_set_ctrl(sgo_world* d, const double* c) { memcpy(d->ctrl, c, sizeof(double) * d->m->nu); }
void sgo_set_dense_solver(sgo_world* d, int on) { d->dense = on; }
void sgo_set_capsule_box_single(sgo_world* d, int on) { d->cb_single = on; }
void sgo_set_implicit_tendon_damping(sgo_world* d, int on) { d->implicit_tendon = on; }
void sgo_set_geom_mask(sgo_world* d, const int* mask) { memcpy(d->geom_mask, mask, sizeof(int) * d->m->ngeom); }
int sgo_status(const sgo_world* d) { return d->status; }
double sgo_last_step_flops(const sgo_world* d) { return d->flops; }
/* op counts accumulated over every mj_forward since the last reset: out[0] = forwards, out[1] = PGS flops (counted per
 * executed row / block update in solve_pgs), out[2] = all stages (PGS + the closed-form count of the other stages of
 * 60 nv + 20 neq + 400 narrowphase pairs + 100 per contact row block + 3000 for the finger chains) */
void sgo_flops_get(const sgo_world* d, double* out) { out[0] = (double)d->steps_total; out[1] = d->flops_pgs_total; out[2] = d->flops_total; }
void sgo_flops_reset(sgo_world* d) { d->steps_total = 0; d->flops_pgs_total = 0; d->flops_total = 0; }

/* mj_resetData: qpos=qpos0, everything else 0 (ref: manenv.py:57 sim.reset()) */
void sgo_reset(sgo_world* d) {
  const sgo_model* m = d->m;
  memcpy(d->qpos, m->qpos0, sizeof(double) * m->nv);
  memset(d->qvel, 0, sizeof(double) * m->nv);
  memset(d->qacc, 0, sizeof(double) * m->nv);
  memset(d->qacc_warmstart, 0, sizeof(double) * m->nv);
  memset(d->act, 0, sizeof(double) * m->nu);
  memset(d->ctrl, 0, sizeof(double) * m->nu);
  memset(d->sensordata, 0, sizeof(double) * m->nsensordata);
  d->time = 0; d->ncon = 0; d->nefc = 0; d->touch_mask = 0;
}

/* ------------------------------------------------------------------------------------------ */
/* small math                                                                                 */
/* ------------------------------------------------------------------------------------------ */
static inline double dot3(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(double* r, const double* a, const double* b) {
  double x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline double norm3(const double* a) { return sqrt(dot3(a, a)); }
static inline double normalize3(double* a) {
  double n = norm3(a);
  if (n < MINVAL) { a[0] = 1; a[1] = 0; a[2] = 0; } else { a[0] /= n; a[1] /= n; a[2] /= n; }
  return n;
}
static void mulquat(double* r, const double* a, const double* b) {
  double t[4] = {a[0] * b[0] - a[1] * b[1] - a[2] * b[2] - a[3] * b[3], a[0] * b[1] + a[1] * b[0] + a[2] * b[3] - a[3] * b[2],
                 a[0] * b[2] - a[1] * b[3] + a[2] * b[0] + a[3] * b[1], a[0] * b[3] + a[1] * b[2] - a[2] * b[1] + a[3] * b[0]};
  memcpy(r, t, sizeof(t));
}
static void quat2mat(double* R, const double* q) {
  double w = q[0], x = q[1], y = q[2], z = q[3];
  R[0] = w * w + x * x - y * y - z * z; R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
  R[3] = 2 * (x * y + w * z); R[4] = w * w - x * x + y * y - z * z; R[5] = 2 * (y * z - w * x);
  R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = w * w - x * x - y * y + z * z;
}
static void mulmatvec3(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[1] * v[1] + R[2] * v[2], y = R[3] * v[0] + R[4] * v[1] + R[5] * v[2],
         z = R[6] * v[0] + R[7] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulmatTvec3(double* r, const double* R, const double* v) {
  double x = R[0] * v[0] + R[3] * v[1] + R[6] * v[2], y = R[1] * v[0] + R[4] * v[1] + R[7] * v[2],
         z = R[2] * v[0] + R[5] * v[1] + R[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static void mulmat3(double* r, const double* A, const double* B) {
  double t[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) t[3 * i + j] = A[3 * i] * B[j] + A[3 * i + 1] * B[3 + j] + A[3 * i + 2] * B[6 + j];
  memcpy(r, t, sizeof(t));
}
static void rotvecquat(double* r, const double* v, const double* q) {
  double R[9]; quat2mat(R, q); mulmatvec3(r, R, v);
}
/* spatial vectors are [rot(3); lin(3)] as in MuJoCo */
static void cross_motion(double* r, const double* vel, const double* v) {
  double a[3], b[3], c[3];
  cross3(a, vel, v); cross3(b, vel, v + 3); cross3(c, vel + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2]; r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static void cross_force(double* r, const double* vel, const double* f) {
  double a[3], b[3], c[3];
  cross3(a, vel, f); cross3(b, vel + 3, f + 3); cross3(c, vel, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2]; r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
static void mul_inert_vec(double* r, const double* i, const double* v) {
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] - i[8] * v[4] + i[7] * v[5];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + i[8] * v[3] - i[6] * v[5];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] - i[7] * v[3] + i[6] * v[4];
  r[3] = i[8] * v[1] - i[7] * v[2] + i[9] * v[3];
  r[4] = i[6] * v[2] - i[8] * v[0] + i[9] * v[4];
  r[5] = i[7] * v[0] - i[6] * v[1] + i[9] * v[5];
}
static inline double dot6(const double* a, const double* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3] + a[4] * b[4] + a[5] * b[5]; }

/* ------------------------------------------------------------------------------------------ */
/* position stage                                                                              */
/* ------------------------------------------------------------------------------------------ */
/* mj_kinematics (SURVEY App. A1 first bullet) */
static void kinematics(sgo_world* d) {
  const sgo_model* m = d->m;
  memset(d->xpos, 0, 3 * sizeof(double));
  d->xquat[0] = 1; d->xquat[1] = d->xquat[2] = d->xquat[3] = 0;
  quat2mat(d->xmat, d->xquat);
  memcpy(d->xipos, d->xpos, 3 * sizeof(double)); memcpy(d->ximat, d->xmat, 9 * sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    double pos[3], quat[4], tmp[3];
    mulmatvec3(pos, d->xmat + 9 * p, d->body_pos + 3 * b);
    for (int k = 0; k < 3; k++) pos[k] += d->xpos[3 * p + k];
    mulquat(quat, d->xquat + 4 * p, m->body_quat + 4 * b);
    for (int j = m->body_jntadr[b]; j < m->body_jntadr[b] + m->body_jntnum[b]; j++) {
      rotvecquat(d->xaxis + 3 * j, m->jnt_axis + 3 * j, quat);
      rotvecquat(tmp, m->jnt_pos + 3 * j, quat);
      for (int k = 0; k < 3; k++) d->xanchor[3 * j + k] = pos[k] + tmp[k];
      double dq = d->qpos[j] - m->qpos0[j];
      if (m->jnt_type[j] == JNT_SLIDE) {
        for (int k = 0; k < 3; k++) pos[k] += d->xaxis[3 * j + k] * dq;
      } else {
        double ql[4] = {1, 0, 0, 0};
        if (dq != 0) { double s = sin(0.5 * dq); ql[0] = cos(0.5 * dq); for (int k = 0; k < 3; k++) ql[1 + k] = m->jnt_axis[3 * j + k] * s; }
        mulquat(quat, quat, ql);
        rotvecquat(tmp, m->jnt_pos + 3 * j, quat);
        for (int k = 0; k < 3; k++) pos[k] = d->xanchor[3 * j + k] - tmp[k];
      }
    }
    double n = sqrt(quat[0] * quat[0] + quat[1] * quat[1] + quat[2] * quat[2] + quat[3] * quat[3]);
    for (int k = 0; k < 4; k++) d->xquat[4 * b + k] = quat[k] / n;
    memcpy(d->xpos + 3 * b, pos, sizeof(pos));
    quat2mat(d->xmat + 9 * b, d->xquat + 4 * b);
    mulmatvec3(tmp, d->xmat + 9 * b, m->body_ipos + 3 * b);
    for (int k = 0; k < 3; k++) d->xipos[3 * b + k] = pos[k] + tmp[k];
    double iq[4]; mulquat(iq, d->xquat + 4 * b, m->body_iquat + 4 * b);
    quat2mat(d->ximat + 9 * b, iq);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g]; double tmp[3], q[4];
    mulmatvec3(tmp, d->xmat + 9 * b, m->geom_pos + 3 * g);
    for (int k = 0; k < 3; k++) d->geom_xpos[3 * g + k] = d->xpos[3 * b + k] + tmp[k];
    mulquat(q, d->xquat + 4 * b, m->geom_quat + 4 * g); quat2mat(d->geom_xmat + 9 * g, q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s]; double tmp[3], q[4];
    mulmatvec3(tmp, d->xmat + 9 * b, m->site_pos + 3 * s);
    for (int k = 0; k < 3; k++) d->site_xpos[3 * s + k] = d->xpos[3 * b + k] + tmp[k];
    mulquat(q, d->xquat + 4 * b, m->site_quat + 4 * s); quat2mat(d->site_xmat + 9 * s, q);
  }
}

/* mj_comPos: subtree coms, com-based inertias (cinert) and motion axes (cdof) */
static void com_pos(sgo_world* d) {
  const sgo_model* m = d->m;
  int nb = m->nbody;
  for (int b = 0; b < nb; b++) for (int k = 0; k < 3; k++) d->subtree_com[3 * b + k] = m->body_mass[b] * d->xipos[3 * b + k];
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; for (int k = 0; k < 3; k++) d->subtree_com[3 * p + k] += d->subtree_com[3 * b + k]; }
  for (int b = 0; b < nb; b++) {
    if (m->body_subtreemass[b] < MINVAL) memcpy(d->subtree_com + 3 * b, d->xipos + 3 * b, 3 * sizeof(double));
    else for (int k = 0; k < 3; k++) d->subtree_com[3 * b + k] /= m->body_subtreemass[b];
  }
  memset(d->cinert, 0, 10 * sizeof(double));
  for (int b = 1; b < nb; b++) {
    const double* R = d->ximat + 9 * b; const double* I = m->body_inertia + 3 * b;
    double dif[3], mass = m->body_mass[b], *res = d->cinert + 10 * b;
    for (int k = 0; k < 3; k++) dif[k] = d->xipos[3 * b + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    /* res_rot = R diag(I) R' */
    res[0] = R[0] * R[0] * I[0] + R[1] * R[1] * I[1] + R[2] * R[2] * I[2];
    res[1] = R[3] * R[3] * I[0] + R[4] * R[4] * I[1] + R[5] * R[5] * I[2];
    res[2] = R[6] * R[6] * I[0] + R[7] * R[7] * I[1] + R[8] * R[8] * I[2];
    res[3] = R[0] * R[3] * I[0] + R[1] * R[4] * I[1] + R[2] * R[5] * I[2];
    res[4] = R[0] * R[6] * I[0] + R[1] * R[7] * I[1] + R[2] * R[8] * I[2];
    res[5] = R[3] * R[6] * I[0] + R[4] * R[7] * I[1] + R[5] * R[8] * I[2];
    res[0] += mass * (dif[1] * dif[1] + dif[2] * dif[2]);
    res[1] += mass * (dif[0] * dif[0] + dif[2] * dif[2]);
    res[2] += mass * (dif[0] * dif[0] + dif[1] * dif[1]);
    res[3] -= mass * dif[0] * dif[1]; res[4] -= mass * dif[0] * dif[2]; res[5] -= mass * dif[1] * dif[2];
    res[6] = mass * dif[0]; res[7] = mass * dif[1]; res[8] = mass * dif[2]; res[9] = mass;
  }
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j]; double off[3], *c = d->cdof + 6 * j;
    for (int k = 0; k < 3; k++) off[k] = d->subtree_com[3 * m->body_rootid[b] + k] - d->xanchor[3 * j + k];
    if (m->jnt_type[j] == JNT_SLIDE) { c[0] = c[1] = c[2] = 0; memcpy(c + 3, d->xaxis + 3 * j, 3 * sizeof(double)); }
    else { memcpy(c, d->xaxis + 3 * j, 3 * sizeof(double)); cross3(c + 3, d->xaxis + 3 * j, off); }
  }
}

/* translational Jacobian column of dof i for a point on its descendants: jacp = cdof_lin + cdof_rot x (p - com_root) */
static inline void jacp_col(const sgo_world* d, int dof, const double* point, double* col) {
  const sgo_model* m = d->m;
  const double* c = d->cdof + 6 * dof; double off[3], t[3];
  int root = m->body_rootid[m->dof_bodyid[dof]];
  for (int k = 0; k < 3; k++) off[k] = point[k] - d->subtree_com[3 * root + k];
  cross3(t, c, off);
  for (int k = 0; k < 3; k++) col[k] = c[3 + k] + t[k];
}

/* mj_tendon (+ mj_transmission for tendon actuators: length = tendon length, moment = gear*ten_J) */
static void tendon(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv;
  memset(d->ten_J, 0, sizeof(double) * m->ntendon * nv);
  for (int t = 0; t < m->ntendon; t++) {
    int adr = m->tendon_adr[t];
    double* J = d->ten_J + (size_t)t * nv;
    if (m->tendon_type[t] == TEN_FIXED) {
      double L = 0;
      for (int w = adr; w < adr + m->tendon_num[t]; w++) { int j = m->wrap_objid[w]; L += m->wrap_prm[w] * d->qpos[j]; J[j] = m->wrap_prm[w]; }
      d->ten_length[t] = L;
    } else {
      int s0 = m->wrap_objid[adr], s1 = m->wrap_objid[adr + 1];
      const double *p0 = d->site_xpos + 3 * s0, *p1 = d->site_xpos + 3 * s1;
      double dv[3] = {p1[0] - p0[0], p1[1] - p0[1], p1[2] - p0[2]};
      double L = normalize3(dv);
      d->ten_length[t] = L;
      for (int side = 0; side < 2; side++) {
        int s = side ? s1 : s0; double sgn = side ? 1.0 : -1.0;
        int b = m->site_bodyid[s];
        while (b > 0) {
          for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++) { double col[3]; jacp_col(d, i, d->site_xpos + 3 * s, col); J[i] += sgn * dot3(dv, col); }
          b = m->body_parentid[b];
        }
      }
    }
  }
}

/* mj_crb + mj_factorM + per-tree dense inverse blocks */
static void solve_ld(const sgo_world* d, double* x) {
  const sgo_model* m = d->m; int nv = m->nv;
  for (int i = nv - 1; i >= 0; i--) if (!m->dof_simple[i] && x[i] != 0) { int adr = m->dof_Madr[i] + 1, j = m->dof_parentid[i]; while (j >= 0) { x[j] -= d->qLD[adr++] * x[i]; j = m->dof_parentid[j]; } }
  for (int i = 0; i < nv; i++) x[i] *= d->qLDiagInv[i];
  for (int i = 0; i < nv; i++) if (!m->dof_simple[i]) { int adr = m->dof_Madr[i] + 1, j = m->dof_parentid[i]; while (j >= 0) { x[i] -= d->qLD[adr++] * x[j]; j = m->dof_parentid[j]; } }
}
static void factor_i(const sgo_model* m, const double* M, double* qLD, double* diaginv) {
  int nv = m->nv;
  memcpy(qLD, M, sizeof(double) * m->nM);
  for (int k = nv - 1; k >= 0; k--) {
    int Mkk = m->dof_Madr[k];
    if (m->dof_simple[k]) continue;
    int Mki = Mkk + 1, i = m->dof_parentid[k];
    while (i >= 0) {
      double tmp = qLD[Mki] / qLD[Mkk];
      int cnt = (i < nv - 1 ? m->dof_Madr[i + 1] : m->nM) - m->dof_Madr[i];
      for (int c = 0; c < cnt; c++) qLD[m->dof_Madr[i] + c] -= qLD[Mki + c] * tmp;
      qLD[Mki] = tmp;
      i = m->dof_parentid[i]; Mki++;
    }
  }
  for (int i = 0; i < nv; i++) diaginv[i] = 1.0 / qLD[m->dof_Madr[i]];
}
static void crb_factor(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv;
  memcpy(d->crb, d->cinert, sizeof(double) * 10 * m->nbody);
  for (int b = m->nbody - 1; b > 0; b--) { int p = m->body_parentid[b]; if (p > 0) for (int k = 0; k < 10; k++) d->crb[10 * p + k] += d->crb[10 * b + k]; }
  memset(d->qM, 0, sizeof(double) * m->nM);
  for (int i = 0; i < nv; i++) {
    double buf[6]; mul_inert_vec(buf, d->crb + 10 * m->dof_bodyid[i], d->cdof + 6 * i);
    int adr = m->dof_Madr[i], j = i;
    while (j >= 0) { d->qM[adr++] = dot6(d->cdof + 6 * j, buf); j = m->dof_parentid[j]; }
  }
  factor_i(m, d->qM, d->qLD, d->qLDiagInv);
  double* x = d->scratch;
  for (int t = 0; t < m->ntree; t++) {
    int n = m->tree_num[t]; const int* dofs = m->tree_dofs + m->tree_adr[t]; double* Mi = d->tree_Minv + m->tree_minvadr[t];
    if (n == 1) { Mi[0] = d->qLDiagInv[dofs[0]]; continue; }
    for (int c = 0; c < n; c++) {
      memset(x, 0, sizeof(double) * nv); x[dofs[c]] = 1; solve_ld(d, x);
      for (int r = 0; r < n; r++) Mi[r * n + c] = x[dofs[r]];
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* collision                                                                                   */
/* ------------------------------------------------------------------------------------------ */
typedef struct { double dist, pos[3], frame[9]; } rawcon;

/* mjraw_SphereBox: normal from sphere (geom1) towards box (geom2) */
static int sphere_box(rawcon* con, double margin, const double* spos, double radius, const double* bpos, const double* bmat, const double* bsize) {
  double tmp[3], center[3], clamped[3], pos[3], nrm[3];
  for (int k = 0; k < 3; k++) tmp[k] = spos[k] - bpos[k];
  mulmatTvec3(center, bmat, tmp);
  for (int k = 0; k < 3; k++) clamped[k] = fmax(-bsize[k], fmin(bsize[k], center[k]));
  for (int k = 0; k < 3; k++) nrm[k] = clamped[k] - center[k];
  double dist = norm3(nrm);
  if (dist - radius > margin) return 0;
  if (dist <= MINVAL) {
    double closest = 2 * (bsize[0] + bsize[1] + bsize[2]); int kbest = 0;
    for (int i = 0; i < 6; i++) {
      double fd = fabs((i % 2 ? 1 : -1) * bsize[i / 2] - center[i / 2]);
      if (fd < closest) { closest = fd; kbest = i; }
    }
    nrm[0] = nrm[1] = nrm[2] = 0; nrm[kbest / 2] = (kbest % 2 ? -1 : 1);
    for (int k = 0; k < 3; k++) pos[k] = center[k] + nrm[k] * (radius - closest) / 2;
    con->dist = -closest - radius;
  } else {
    for (int k = 0; k < 3; k++) nrm[k] /= dist;
    for (int k = 0; k < 3; k++) pos[k] = 0.5 * (clamped[k] + center[k] + nrm[k] * radius);
    con->dist = dist - radius;
  }
  mulmatvec3(con->frame, bmat, nrm);
  mulmatvec3(tmp, bmat, pos);
  for (int k = 0; k < 3; k++) con->pos[k] = tmp[k] + bpos[k];
  for (int k = 3; k < 9; k++) con->frame[k] = 0;
  return 1;
}

/* derivative (up to a factor 2) of the squared distance of p(t)=c+t*h to the box, and the distance itself */
static double seg_box_grad(const double* c, const double* h, const double* s, double t, double* d2) {
  double g = 0, q = 0;
  for (int k = 0; k < 3; k++) {
    double p = c[k] + t * h[k];
    double e = p - fmax(-s[k], fmin(s[k], p));
    g += h[k] * e; q += e * e;
  }
  if (d2) *d2 = q;
  return g;
}

/* capsule-box, DEFINED HERE (see file header): contact at the segment point of minimal signed
 * distance to the box; a second contact at the far end of the segment if that end is within margin. */
static int capsule_box_mode(rawcon* con, double margin, const double* cpos, const double* cmat, const double* csize,
                            const double* bpos, const double* bmat, const double* bsize, int second);
static int capsule_box(rawcon* con, double margin, const double* cpos, const double* cmat, const double* csize,
                       const double* bpos, const double* bmat, const double* bsize) {
  return capsule_box_mode(con, margin, cpos, cmat, csize, bpos, bmat, bsize, 1);
}
/* second = 0 drops the second contact (sensitivity study of the one rule of this narrowphase that is a choice rather than
 * geometry, tests/test_oracle.py::test_second_capsule_box_contact_is_immaterial); the product and every parity test use 1 */
static int capsule_box_mode(rawcon* con, double margin, const double* cpos, const double* cmat, const double* csize,
                            const double* bpos, const double* bmat, const double* bsize, int second) {
  double radius = csize[0], hl = csize[1];
  double axis_w[3] = {cmat[2], cmat[5], cmat[8]};
  double tmp[3], c[3], ax[3], h[3];
  for (int k = 0; k < 3; k++) tmp[k] = cpos[k] - bpos[k];
  mulmatTvec3(c, bmat, tmp);
  mulmatTvec3(ax, bmat, axis_w);
  for (int k = 0; k < 3; k++) h[k] = ax[k] * hl;
  /* root of the monotone piecewise-linear derivative g(t) on [-1,1] */
  double tlo = -1, thi = 1, d2;
  double glo = seg_box_grad(c, h, bsize, -1, NULL), ghi = seg_box_grad(c, h, bsize, 1, NULL);
  double tstar;
  if (glo >= 0) tstar = -1;
  else if (ghi <= 0) tstar = 1;
  else {
    for (int k = 0; k < 3; k++) {
      if (fabs(h[k]) < MINVAL) continue;
      for (int sgn = -1; sgn <= 1; sgn += 2) {
        double t = (sgn * bsize[k] - c[k]) / h[k];
        if (t <= tlo || t >= thi) continue;
        double g = seg_box_grad(c, h, bsize, t, NULL);
        if (g <= 0) { tlo = t; glo = g; } else { thi = t; ghi = g; }
      }
    }
    tstar = (ghi - glo > MINVAL) ? tlo + (0 - glo) * (thi - tlo) / (ghi - glo) : tlo;
  }
  seg_box_grad(c, h, bsize, tstar, &d2);
  if (d2 <= MINVAL * MINVAL) {
    /* the segment enters the box: use the end that is inside (the deeper one if both are), or the middle of
     * the inside interval when the segment passes through */
    double t0 = -1, t1 = 1;
    for (int k = 0; k < 3; k++) {
      if (fabs(h[k]) < MINVAL) continue;
      double ta = (-bsize[k] - c[k]) / h[k], tb = (bsize[k] - c[k]) / h[k];
      if (ta > tb) { double x = ta; ta = tb; tb = x; }
      if (ta > t0) t0 = ta;
      if (tb < t1) t1 = tb;
    }
    /* the six planes: depth_i(t) = a_i + b_i t ; depth(t) = min_i depth_i(t) */
    double a[6], b[6];
    for (int k = 0; k < 3; k++) { a[2 * k] = bsize[k] - c[k]; b[2 * k] = -h[k]; a[2 * k + 1] = bsize[k] + c[k]; b[2 * k + 1] = h[k]; }
    int in0 = (t0 <= -1), in1 = (t1 >= 1);
    if (in0 && in1) {
      double dm = 1e300, dp = 1e300;
      for (int q = 0; q < 6; q++) { dm = fmin(dm, a[q] - b[q]); dp = fmin(dp, a[q] + b[q]); }
      tstar = (dp > dm) ? 1 : -1;
    } else if (in0) tstar = -1;
    else if (in1) tstar = 1;
    else tstar = 0.5 * (t0 + t1);
  }
  int n = 0; double sp[3];
  for (int k = 0; k < 3; k++) sp[k] = cpos[k] + axis_w[k] * (tstar * hl);
  n += sphere_box(con + n, margin, sp, radius, bpos, bmat, bsize);
  if (!second) return n;
  double t2 = (tstar >= 0) ? -1.0 : 1.0;
  for (int k = 0; k < 3; k++) sp[k] = cpos[k] + axis_w[k] * (t2 * hl);
  n += sphere_box(con + n, margin, sp, radius, bpos, bmat, bsize);
  return n;
}

/* mjc_PlaneCapsule: a sphere-plane test at both ends; tangent hint = capsule axis */
static int plane_capsule(rawcon* con, double margin, const double* ppos, const double* pmat, const double* cpos, const double* cmat, const double* csize) {
  double nrm[3] = {pmat[2], pmat[5], pmat[8]}, axis[3] = {cmat[2], cmat[5], cmat[8]};
  int n = 0;
  for (int side = 1; side >= -1; side -= 2) {
    double sp[3], dif[3];
    for (int k = 0; k < 3; k++) { sp[k] = cpos[k] + side * axis[k] * csize[1]; dif[k] = sp[k] - ppos[k]; }
    double dist = dot3(dif, nrm) - csize[0];
    if (dist > margin) continue;
    con[n].dist = dist;
    for (int k = 0; k < 3; k++) { con[n].pos[k] = sp[k] - nrm[k] * (csize[0] + 0.5 * dist); con[n].frame[k] = nrm[k]; con[n].frame[3 + k] = axis[k]; con[n].frame[6 + k] = 0; }
    n++;
  }
  return n;
}

/* separating-axis overlap test for two boxes (detection only: box-box contacts are not restated) */
static int box_box_overlap(const double* p1, const double* R1, const double* s1, const double* p2, const double* R2, const double* s2) {
  double Rr[9], T[3], tmp[3], A[9];
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) { Rr[3 * i + j] = R1[i] * R2[j] + R1[3 + i] * R2[3 + j] + R1[6 + i] * R2[6 + j]; A[3 * i + j] = fabs(Rr[3 * i + j]) + 1e-12; }
  for (int k = 0; k < 3; k++) tmp[k] = p2[k] - p1[k];
  mulmatTvec3(T, R1, tmp);
  for (int i = 0; i < 3; i++) if (fabs(T[i]) > s1[i] + s2[0] * A[3 * i] + s2[1] * A[3 * i + 1] + s2[2] * A[3 * i + 2]) return 0;
  for (int j = 0; j < 3; j++) if (fabs(T[0] * Rr[j] + T[1] * Rr[3 + j] + T[2] * Rr[6 + j]) > s2[j] + s1[0] * A[j] + s1[1] * A[3 + j] + s1[2] * A[6 + j]) return 0;
  for (int i = 0; i < 3; i++) for (int j = 0; j < 3; j++) {
    int i1 = (i + 1) % 3, i2 = (i + 2) % 3, j1 = (j + 1) % 3, j2 = (j + 2) % 3;
    double ra = s1[i1] * A[3 * i2 + j] + s1[i2] * A[3 * i1 + j];
    double rb = s2[j1] * A[3 * i + j2] + s2[j2] * A[3 * i + j1];
    if (fabs(T[i2] * Rr[3 * i1 + j] - T[i1] * Rr[3 * i2 + j]) > ra + rb) return 0;
  }
  return 1;
}

/* mju_makeFrame */
static void make_frame(double* f) {
  normalize3(f);
  if (norm3(f + 3) < 0.5) { f[3] = f[4] = f[5] = 0; if (f[1] < 0.5 && f[1] > -0.5) f[4] = 1; else f[5] = 1; }
  double dp = dot3(f, f + 3);
  for (int k = 0; k < 3; k++) f[3 + k] -= f[k] * dp;
  normalize3(f + 3);
  cross3(f + 6, f, f + 3);
}

/* mj_collision over the statically filtered pair list */
static void collision(sgo_world* d) {
  const sgo_model* m = d->m;
  d->ncon = 0; d->touch_mask = 0; d->ncand = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_g1[p], g2 = m->pair_g2[p];
    int t1 = m->geom_type[g1], t2 = m->geom_type[g2];
    double margin = fmax(m->geom_margin[g1], m->geom_margin[g2]);
    double gap = fmax(m->geom_gap[g1], m->geom_gap[g2]);
    const double *p1 = d->geom_xpos + 3 * g1, *p2 = d->geom_xpos + 3 * g2, *R1 = d->geom_xmat + 9 * g1, *R2 = d->geom_xmat + 9 * g2;
    /* bounding-sphere cull (plane: signed distance of the centre) */
    if (t1 == GEOM_PLANE) {
      double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]}, nrm[3] = {R1[2], R1[5], R1[8]};
      if (dot3(dif, nrm) > margin + m->geom_rbound[g2]) continue;
    } else {
      double dif[3] = {p2[0] - p1[0], p2[1] - p1[1], p2[2] - p1[2]};
      double bound = m->geom_rbound[g1] + m->geom_rbound[g2] + margin;
      if (dot3(dif, dif) > bound * bound) continue;
    }
    rawcon rc[4]; int n = 0;
    d->ncand++;
    if (t1 == GEOM_PLANE && t2 == GEOM_CAPSULE) n = plane_capsule(rc, margin, p1, R1, p2, R2, m->geom_size + 3 * g2);
    else if (t1 == GEOM_SPHERE && t2 == GEOM_BOX) n = sphere_box(rc, margin, p1, m->geom_size[3 * g1], p2, R2, m->geom_size + 3 * g2);
    else if (t1 == GEOM_CAPSULE && t2 == GEOM_BOX) n = capsule_box_mode(rc, margin, p1, R1, m->geom_size + 3 * g1, p2, R2, m->geom_size + 3 * g2, !d->cb_single);
    else if (t1 == GEOM_BOX && t2 == GEOM_BOX) { if (box_box_overlap(p1, R1, m->geom_size + 3 * g1, p2, R2, m->geom_size + 3 * g2)) d->step_status |= SGO_ST_BOXBOX; }
    else d->step_status |= SGO_ST_BOXBOX; /* unsupported pair reached the narrowphase (e.g. plane-box) */
    for (int i = 0; i < n; i++) {
      if (d->ncon >= m->nconmax) { d->step_status |= SGO_ST_CON_FULL; break; }
      int c = d->ncon++;
      d->con_dist[c] = rc[i].dist;
      memcpy(d->con_pos + 3 * c, rc[i].pos, 3 * sizeof(double));
      memcpy(d->con_frame + 9 * c, rc[i].frame, 9 * sizeof(double));
      make_frame(d->con_frame + 9 * c);
      d->con_geom1[c] = g1; d->con_geom2[c] = g2;
      /* mj_contactParam: condim max, friction max, solref/solimp mixed by solmix */
      d->con_dim[c] = m->geom_condim[g1] > m->geom_condim[g2] ? m->geom_condim[g1] : m->geom_condim[g2];
      double f[3];
      for (int k = 0; k < 3; k++) f[k] = fmax(m->geom_friction[3 * g1 + k], m->geom_friction[3 * g2 + k]);
      double* fr = d->con_friction + 5 * c; fr[0] = f[0]; fr[1] = f[0]; fr[2] = f[1]; fr[3] = f[2]; fr[4] = f[2];
      double mix = m->geom_solmix[g1] / (m->geom_solmix[g1] + m->geom_solmix[g2]);
      for (int k = 0; k < 2; k++) d->con_solref[2 * c + k] = mix * m->geom_solref[2 * g1 + k] + (1 - mix) * m->geom_solref[2 * g2 + k];
      for (int k = 0; k < 5; k++) d->con_solimp[5 * c + k] = mix * m->geom_solimp[5 * g1 + k] + (1 - mix) * m->geom_solimp[5 * g2 + k];
      d->con_includemargin[c] = margin - gap;
      d->con_exclude[c] = (rc[i].dist >= margin - gap);
      d->con_efc[c] = -1;
      int mk = d->geom_mask[g1] | d->geom_mask[g2];
      d->touch_mask |= (1 << 30);
      if (mk & 1) d->touch_mask |= (mk >> 1);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* constraints                                                                                 */
/* ------------------------------------------------------------------------------------------ */
/* append one row given (col,val) pairs; the pattern is widened to whole dof-trees so that
 * B = M^-1 J^T lives on the same pattern */
static int add_row(sgo_world* d, int type, int id, double pos, double margin, int n, const int* cols, const double* vals) {
  const sgo_model* m = d->m;
  if (d->nefc >= m->njmax) { d->step_status |= SGO_ST_EFC_FULL; return -1; }
  int r = d->nefc++;
  d->efc_type[r] = type; d->efc_id[r] = id; d->efc_pos[r] = pos; d->efc_margin[r] = margin;
  int adr = d->nnz, cnt = 0;
  int* ci = d->efc_colind + adr;
  for (int k = 0; k < n; k++) {
    int c = cols[k];
    if (m->dof_simple[c]) { if (!d->mark[c]) { d->mark[c] = 1; ci[cnt++] = c; } }
    else { int t = m->dof_treeid[c]; for (int q = 0; q < m->tree_num[t]; q++) { int cc = m->tree_dofs[m->tree_adr[t] + q]; if (!d->mark[cc]) { d->mark[cc] = 1; ci[cnt++] = cc; } } }
  }
  /* sort columns ascending (insertion sort; rows are short except the tendon row, already sorted) */
  for (int a = 1; a < cnt; a++) { int v = ci[a], b = a - 1; while (b >= 0 && ci[b] > v) { ci[b + 1] = ci[b]; b--; } ci[b + 1] = v; }
  double* x = d->scratch;     /* dense scatter, only touched entries are used */
  for (int k = 0; k < cnt; k++) { x[ci[k]] = 0; d->mark[ci[k]] = 0; }
  for (int k = 0; k < n; k++) x[cols[k]] += vals[k];
  for (int k = 0; k < cnt; k++) d->efc_J[adr + k] = x[ci[k]];
  /* B = M^-1 J^T, tree block by tree block */
  for (int k = 0; k < cnt; k++) {
    int c = ci[k], t = m->dof_treeid[c], nt = m->tree_num[t];
    if (nt == 1) { d->efc_B[adr + k] = d->tree_Minv[m->tree_minvadr[t]] * x[c]; continue; }
    const int* dofs = m->tree_dofs + m->tree_adr[t]; const double* Mi = d->tree_Minv + m->tree_minvadr[t];
    int rloc = 0; while (dofs[rloc] != c) rloc++;
    double s = 0;
    for (int q = 0; q < nt; q++) s += Mi[rloc * nt + q] * x[dofs[q]];
    d->efc_B[adr + k] = s;
  }
  d->efc_rowadr[r] = adr; d->efc_rownnz[r] = cnt; d->nnz += cnt;
  return r;
}

/* mj_makeConstraint: equality rows in id order, then joint limits, then elliptic contacts */
static void make_constraint(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv;
  d->nefc = d->ne = d->nl = d->nnz = 0;
  int* cols = (int*)malloc(sizeof(int) * (nv + 64));
  double* vals = d->scratch + 2 * nv;
  /* 1. equalities (mj_instantiateEquality, joint/tendon polynomial couplings) */
  for (int e = 0; e < m->neq; e++) {
    int o1 = m->eq_obj1id[e], o2 = m->eq_obj2id[e]; const double* dat = m->eq_data + 5 * e;
    int n = 0; double pos;
    if (m->eq_type[e] == EQ_JOINT) {
      double p1 = d->qpos[o1] - m->qpos0[o1];
      cols[n] = o1; vals[n++] = 1;
      if (o2 >= 0) {
        double dif = d->qpos[o2] - m->qpos0[o2];
        pos = p1 - dat[0] - (dat[1] * dif + dat[2] * dif * dif + dat[3] * dif * dif * dif + dat[4] * dif * dif * dif * dif);
        double deriv = dat[1] + 2 * dat[2] * dif + 3 * dat[3] * dif * dif + 4 * dat[4] * dif * dif * dif;
        cols[n] = o2; vals[n++] = -deriv;
      } else pos = p1 - dat[0];
    } else {
      double p1 = d->ten_length[o1] - m->tendon_length0[o1];
      const double* J1 = d->ten_J + (size_t)o1 * nv;
      double deriv = 0, dif = 0;
      if (o2 >= 0) {
        dif = d->ten_length[o2] - m->tendon_length0[o2];
        pos = p1 - dat[0] - (dat[1] * dif + dat[2] * dif * dif + dat[3] * dif * dif * dif + dat[4] * dif * dif * dif * dif);
        deriv = dat[1] + 2 * dat[2] * dif + 3 * dat[3] * dif * dif + 4 * dat[4] * dif * dif * dif;
      } else pos = p1 - dat[0];
      for (int i = 0; i < nv; i++) {
        double v = J1[i] - (o2 >= 0 ? deriv * d->ten_J[(size_t)o2 * nv + i] : 0);
        if (v != 0) { cols[n] = i; vals[n++] = v; }
      }
    }
    add_row(d, CNSTR_EQUALITY, e, pos, 0, n, cols, vals);
  }
  d->ne = d->nefc;
  /* 3. joint limits (mj_instantiateLimit): lower then upper */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j]) continue;
    for (int side = -1; side <= 1; side += 2) {
      double dist = side * (m->jnt_range[2 * j + (side + 1) / 2] - d->qpos[j]);
      if (dist < m->jnt_margin[j]) { cols[0] = j; vals[0] = -side; add_row(d, CNSTR_LIMIT_JOINT, j, dist, m->jnt_margin[j], 1, cols, vals); }
    }
  }
  d->nl = d->nefc - d->ne;
  /* 4. contacts (mj_instantiateContact, elliptic cones): J = frame * (jacp(body2) - jacp(body1)) */
  for (int c = 0; c < d->ncon; c++) {
    if (d->con_exclude[c]) continue;
    int dim = d->con_dim[c];
    if (dim != 3) { d->step_status |= SGO_ST_BOXBOX; continue; }
    int bodies[2] = {m->geom_bodyid[d->con_geom1[c]], m->geom_bodyid[d->con_geom2[c]]};
    int n = 0; double jac[3][64];
    for (int s = 0; s < 2; s++) {
      int b = bodies[s]; double sgn = s ? 1.0 : -1.0;
      while (b > 0) {
        for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++) {
          double col[3]; jacp_col(d, i, d->con_pos + 3 * c, col);
          if (n < 64) { cols[n] = i; for (int k = 0; k < 3; k++) jac[k][n] = sgn * dot3(d->con_frame + 9 * c + 3 * k, col); n++; }
        }
        b = m->body_parentid[b];
      }
    }
    for (int k = 0; k < dim; k++) {
      int r = add_row(d, CNSTR_CONTACT_ELLIPTIC, c, k == 0 ? d->con_dist[c] : 0, k == 0 ? d->con_includemargin[c] : 0, n, cols, jac[k]);
      if (k == 0) d->con_efc[c] = r;
    }
  }
  free(cols);
}

/* mj_makeImpedance (getsolparam + getimpedance + diagApprox + R, K, B), SURVEY App. A1 last bullets */
static double impedance(const double* si_in, double pos, double margin) {
  double si[5] = {fmin(MAXIMP, fmax(MINIMP, si_in[0])), fmin(MAXIMP, fmax(MINIMP, si_in[1])), fmax(0, si_in[2]),
                  fmin(MAXIMP, fmax(MINIMP, si_in[3])), fmax(1, si_in[4])};
  if (si[0] == si[1] || si[2] <= MINVAL) return 0.5 * (si[0] + si[1]);
  double x = fabs((pos - margin) / si[2]);
  if (x >= 1) return si[1];
  if (x <= 0) return si[0];
  double y;
  if (si[4] == 1) y = x;
  else if (x <= si[3]) y = pow(x, si[4]) / pow(si[3], si[4] - 1);
  else y = 1 - pow(1 - x, si[4]) / pow(1 - si[3], si[4] - 1);
  return si[0] + y * (si[1] - si[0]);
}
static void make_impedance(sgo_world* d) {
  const sgo_model* m = d->m;
  for (int i = 0; i < d->nefc; i++) {
    int id = d->efc_id[i]; const double *solref, *solimp; double diag;
    switch (d->efc_type[i]) {
      case CNSTR_EQUALITY:
        solref = m->eq_solref + 2 * id; solimp = m->eq_solimp + 5 * id;
        if (m->eq_type[id] == EQ_JOINT) { diag = m->dof_invweight0[m->eq_obj1id[id]]; if (m->eq_obj2id[id] >= 0) diag += m->dof_invweight0[m->eq_obj2id[id]]; }
        else { diag = m->tendon_invweight0[m->eq_obj1id[id]]; if (m->eq_obj2id[id] >= 0) diag += m->tendon_invweight0[m->eq_obj2id[id]]; }
        break;
      case CNSTR_LIMIT_JOINT:
        solref = m->jnt_solref + 2 * id; solimp = m->jnt_solimp + 5 * id; diag = m->dof_invweight0[id];
        break;
      default: {
        solref = d->con_solref + 2 * id; solimp = d->con_solimp + 5 * id;
        int b1 = m->geom_bodyid[d->con_geom1[id]], b2 = m->geom_bodyid[d->con_geom2[id]];
        diag = m->body_invweight0[2 * b1] + m->body_invweight0[2 * b2];   /* translational; dim 3 has no rotational rows */
      }
    }
    /* friction rows of an elliptic contact copy the normal row's impedance */
    int first = i;
    if (d->efc_type[i] == CNSTR_CONTACT_ELLIPTIC) first = d->con_efc[id];
    double imp = (first == i) ? impedance(solimp, d->efc_pos[i], d->efc_margin[i]) : d->efc_imp[first];
    double dmax = fmin(MAXIMP, fmax(MINIMP, solimp[1]));
    double K, B;
    if (solref[0] > 0) {
      double tc = fmax(solref[0], 2 * m->timestep);      /* refsafe */
      K = 1 / fmax(MINVAL, dmax * dmax * tc * tc * solref[1] * solref[1]);
      B = 2 / fmax(MINVAL, dmax * tc);
    } else { K = -solref[0] / fmax(MINVAL, dmax * dmax); B = -solref[1] / fmax(MINVAL, dmax); }
    if (first != i) K = 0;
    d->efc_diagApprox[i] = diag; d->efc_imp[i] = imp; d->efc_K[i] = K; d->efc_Bd[i] = B;
    d->efc_R[i] = fmax(MINVAL, (1 - imp) * diag / imp);
  }
  /* elliptic cones: R[1] = R[0]/impratio, mu = friction[0]*sqrt(R[1]/R[0]), R[j] = R[1]*mu0^2/mu_{j-1}^2 */
  for (int c = 0; c < d->ncon; c++) {
    int i = d->con_efc[c]; if (i < 0) continue;
    const double* fr = d->con_friction + 5 * c;
    d->efc_R[i + 1] = d->efc_R[i] / fmax(MINVAL, m->impratio);
    d->con_mu[c] = fr[0] * sqrt(d->efc_R[i + 1] / d->efc_R[i]);
    for (int j = 2; j < d->con_dim[c]; j++) d->efc_R[i + j] = d->efc_R[i + 1] * fr[0] * fr[0] / (fr[j - 1] * fr[j - 1]);
  }
  for (int i = 0; i < d->nefc; i++) d->efc_D[i] = 1 / d->efc_R[i];
}

static inline double row_dot(const sgo_world* d, int r, const double* x) {
  double s = 0; int adr = d->efc_rowadr[r];
  for (int k = 0; k < d->efc_rownnz[r]; k++) s += d->efc_J[adr + k] * x[d->efc_colind[adr + k]];
  return s;
}

/* ------------------------------------------------------------------------------------------ */
/* velocity stage, forces                                                                      */
/* ------------------------------------------------------------------------------------------ */
/* mj_comVel */
static void com_vel(sgo_world* d) {
  const sgo_model* m = d->m;
  memset(d->cvel, 0, 6 * sizeof(double));
  for (int b = 1; b < m->nbody; b++) {
    double cv[6]; memcpy(cv, d->cvel + 6 * m->body_parentid[b], sizeof(cv));
    for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++) {
      cross_motion(d->cdof_dot + 6 * i, cv, d->cdof + 6 * i);
      for (int k = 0; k < 6; k++) cv[k] += d->cdof[6 * i + k] * d->qvel[i];
    }
    memcpy(d->cvel + 6 * b, cv, sizeof(cv));
  }
}
/* mj_passive: joint and tendon spring-dampers */
static void passive(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv;
  for (int i = 0; i < nv; i++) d->qfrc_passive[i] = -d->jnt_stiffness[i] * (d->qpos[i] - m->qpos_spring[i]) - d->dof_damping[i] * d->qvel[i];
  for (int t = 0; t < m->ntendon; t++) {
    double k = d->tendon_stiffness[t], b = d->tendon_damping[t];
    if (k == 0 && b == 0) continue;
    double frc = -k * (d->ten_length[t] - m->tendon_lengthspring[t]) - b * d->ten_velocity[t];
    for (int i = 0; i < nv; i++) d->qfrc_passive[i] += d->ten_J[(size_t)t * nv + i] * frc;
  }
}
/* mj_rne with qacc = 0 (bias: gravity + Coriolis/centrifugal); with_acc: mj_rnePostConstraint's cacc */
static void rne(sgo_world* d, int with_acc, double* result) {
  const sgo_model* m = d->m; int nb = m->nbody;
  double* cfrc = d->cfrc;
  double* cacc = d->cacc;
  cacc[0] = cacc[1] = cacc[2] = 0; for (int k = 0; k < 3; k++) cacc[3 + k] = -m->gravity[k];
  memset(cfrc, 0, 6 * sizeof(double));
  for (int b = 1; b < nb; b++) {
    double* a = cacc + 6 * b; memcpy(a, cacc + 6 * m->body_parentid[b], 6 * sizeof(double));
    for (int i = m->body_dofadr[b]; i < m->body_dofadr[b] + m->body_dofnum[b]; i++) {
      for (int k = 0; k < 6; k++) a[k] += d->cdof_dot[6 * i + k] * d->qvel[i];
      if (with_acc) for (int k = 0; k < 6; k++) a[k] += d->cdof[6 * i + k] * d->qacc[i];
    }
    if (!result) continue;
    double t1[6], t2[6];
    mul_inert_vec(cfrc + 6 * b, d->cinert + 10 * b, a);
    mul_inert_vec(t1, d->cinert + 10 * b, d->cvel + 6 * b);
    cross_force(t2, d->cvel + 6 * b, t1);
    for (int k = 0; k < 6; k++) cfrc[6 * b + k] += t2[k];
  }
  if (!result) return;
  for (int b = nb - 1; b > 0; b--) { int p = m->body_parentid[b]; if (p > 0) for (int k = 0; k < 6; k++) cfrc[6 * p + k] += cfrc[6 * b + k]; }
  for (int i = 0; i < m->nv; i++) result[i] = dot6(d->cdof + 6 * i, cfrc + 6 * m->dof_bodyid[i]);
}

/* mj_sensorVel (gyro) / mj_sensorAcc (accelerometer), SURVEY App. A5 */
static void sensors(sgo_world* d, int stage) {
  const sgo_model* m = d->m;
  for (int s = 0; s < m->nsensor; s++) {
    int site = m->sensor_objid[s], b = m->site_bodyid[site];
    const double* R = d->site_xmat + 9 * site;
    double* out = d->sensordata + m->sensor_adr[s];
    double off[3];
    for (int k = 0; k < 3; k++) off[k] = d->site_xpos[3 * site + k] - d->subtree_com[3 * m->body_rootid[b] + k];
    const double* cv = d->cvel + 6 * b;
    if (stage == 0 && m->sensor_type[s] == SENS_GYRO) mulmatTvec3(out, R, cv);
    if (stage == 1 && m->sensor_type[s] == SENS_ACCEL) {
      const double* ca = d->cacc + 6 * b;
      double t[3], vlin[3], alin[3], corr[3];
      cross3(t, cv, off); for (int k = 0; k < 3; k++) vlin[k] = cv[3 + k] + t[k];
      cross3(t, ca, off); for (int k = 0; k < 3; k++) alin[k] = ca[3 + k] + t[k];
      cross3(corr, cv, vlin);
      for (int k = 0; k < 3; k++) alin[k] += corr[k];
      mulmatTvec3(out, R, alin);
    }
  }
}

/* ------------------------------------------------------------------------------------------ */
/* constraint solver                                                                           */
/* ------------------------------------------------------------------------------------------ */
/* mj_constraintUpdate restricted to the force part (used by the warm start) */
static void constraint_update(sgo_world* d, const double* jar) {
  for (int i = 0; i < d->nefc; i++) d->efc_force[i] = -d->efc_D[i] * jar[i];
  for (int i = d->ne; i < d->nefc; i++) {
    if (d->efc_type[i] != CNSTR_CONTACT_ELLIPTIC) { if (jar[i] >= 0) d->efc_force[i] = 0; continue; }
    int c = d->efc_id[i], dim = d->con_dim[c]; const double* fr = d->con_friction + 5 * c; double mu = d->con_mu[c];
    double U[6]; U[0] = jar[i] * mu; double T = 0;
    for (int j = 1; j < dim; j++) { U[j] = jar[i + j] * fr[j - 1]; T += U[j] * U[j]; }
    T = sqrt(T); double N = U[0];
    if (N >= mu * T || (T <= 0 && N >= 0)) { for (int j = 0; j < dim; j++) d->efc_force[i + j] = 0; }
    else if (mu * N + T <= 0 || (T <= 0 && N < 0)) { /* bottom zone: quadratic, forces stay -D*jar */ }
    else {
      double Dm = d->efc_D[i] / (mu * mu * (1 + mu * mu)), NT = N - mu * T;
      d->efc_force[i] = -Dm * NT * mu;
      for (int j = 1; j < dim; j++) d->efc_force[i + j] = -d->efc_force[i] / T * U[j] * fr[j - 1];
    }
    i += dim - 1;
  }
}

/* mju_QCQP2 */
static int qcqp2(double* res, const double* Ain, const double* bin, const double* dd, double r) {
  double b1 = bin[0] * dd[0], b2 = bin[1] * dd[1];
  double A11 = Ain[0] * dd[0] * dd[0], A22 = Ain[3] * dd[1] * dd[1], A12 = Ain[1] * dd[0] * dd[1];
  double la = 0, v1 = 0, v2 = 0;
  for (int iter = 0; iter < 20; iter++) {
    double det = (A11 + la) * (A22 + la) - A12 * A12;
    if (det < 1e-10) { res[0] = 0; res[1] = 0; return 0; }
    double detinv = 1 / det, P11 = (A22 + la) * detinv, P22 = (A11 + la) * detinv, P12 = -A12 * detinv;
    v1 = -P11 * b1 - P12 * b2; v2 = -P12 * b1 - P22 * b2;
    double val = v1 * v1 + v2 * v2 - r * r;
    if (val < 1e-10) break;
    double deriv = -2 * (P11 * v1 * v1 + 2 * P12 * v1 * v2 + P22 * v2 * v2);
    double delta = -val / deriv;
    if (delta < 1e-10) break;
    la += delta;
  }
  res[0] = v1 * dd[0]; res[1] = v2 * dd[1];
  return la != 0;
}

static double cost_change(const double* A, double* force, const double* old, const double* res, int dim) {
  double change;
  if (dim == 1) { double dl = force[0] - old[0]; change = 0.5 * dl * dl * A[0] + dl * res[0]; }
  else {
    double dl[6]; change = 0;
    for (int j = 0; j < dim; j++) dl[j] = force[j] - old[j];
    for (int j = 0; j < dim; j++) { double s = 0; for (int k = 0; k < dim; k++) s += A[j * dim + k] * dl[k]; change += 0.5 * dl[j] * s + dl[j] * res[j]; }
  }
  if (change > 1e-10) { memcpy(force, old, sizeof(double) * dim); change = 0; }
  return change;
}

/* one PGS update of the row/block starting at i given its residual and diagonal block (mj_solPGS inner body).
 * Returns the cost change (<= 0). */
static double pgs_block(sgo_world* d, int i, int dim, const double* res, const double* Athis) {
  double* force = d->efc_force; double old[6];
  memcpy(old, force + i, sizeof(double) * dim);
  if (dim == 1) {
    force[i] -= res[0] / Athis[0];
    if (i >= d->ne && force[i] < 0) force[i] = 0;
  } else {
    int c = d->efc_id[i]; const double* fr = d->con_friction + 5 * c;
    if (force[i] < MINVAL) {
      force[i] -= res[0] / Athis[0];
      if (force[i] < 0) force[i] = 0;
      for (int j = 1; j < dim; j++) force[i + j] = 0;
    } else {
      double v[6], v1[6], denom = 0;
      memcpy(v, force + i, sizeof(double) * dim);
      for (int j = 0; j < dim; j++) { double s = 0; for (int k = 0; k < dim; k++) s += Athis[j * dim + k] * v[k]; v1[j] = s; denom += v[j] * s; }
      if (denom >= MINVAL) {
        double x = 0; for (int j = 0; j < dim; j++) x += v[j] * res[j];
        x = -x / denom;
        if (force[i] + x * v[0] < 0) x = -v[0] / v[0];
        for (int j = 0; j < dim; j++) force[i + j] += x * v[j];
      }
    }
    double Ac[25], bc[5], vv[5];
    for (int j = 0; j < dim - 1; j++) {
      bc[j] = res[j + 1];
      for (int k = 0; k < dim - 1; k++) { Ac[j * (dim - 1) + k] = Athis[(j + 1) * dim + (k + 1)]; bc[j] -= Ac[j * (dim - 1) + k] * old[1 + k]; }
      bc[j] += Athis[(j + 1) * dim] * (force[i] - old[0]);
    }
    if (force[i] < MINVAL) for (int j = 1; j < dim; j++) force[i + j] = 0;
    else {
      int active = qcqp2(vv, Ac, bc, fr, force[i]);
      if (active) {
        double s = 0; for (int j = 0; j < dim - 1; j++) s += vv[j] * vv[j] / (fr[j] * fr[j]);
        s = sqrt(force[i] * force[i] / fmax(MINVAL, s));
        for (int j = 0; j < dim - 1; j++) vv[j] *= s;
      }
      for (int j = 0; j < dim - 1; j++) force[i + 1 + j] = vv[j];
    }
  }
  return cost_change(Athis, force + i, old, res, dim);
}

/* matrix-free PGS: qacc = qacc_smooth + M^-1 J^T f is kept up to date, res_i = J_i qacc - aref_i + R_i f_i */
static void solve_pgs(sgo_world* d) {
  const sgo_model* m = d->m; int nefc = d->nefc;
  double scale = 1 / (m->meaninertia * (m->nv > 1 ? m->nv : 1));
  double* qacc = d->qacc; double flops = 0;
  /* diagonal blocks of AR = J M^-1 J' + R, once per step */
  for (int i = 0; i < nefc; i++) {
    int dim = (d->efc_type[i] == CNSTR_CONTACT_ELLIPTIC) ? d->con_dim[d->efc_id[i]] : 1;
    for (int j = 0; j < dim; j++) {
      int r = i + j, adr = d->efc_rowadr[r], nn = d->efc_rownnz[r];
      for (int k = 0; k < dim; k++) {   /* rows of one contact share a pattern */
        int adk = d->efc_rowadr[i + k]; double s = 0;
        for (int q = 0; q < nn; q++) s += d->efc_J[adr + q] * d->efc_B[adk + q];
        d->efc_A[3 * r + k] = s + (j == k ? d->efc_R[r] : 0);
      }
      flops += 2.0 * nn * dim;
    }
    i += dim - 1;
  }
  int iter = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int i = 0; i < nefc; i++) {
      int dim = (d->efc_type[i] == CNSTR_CONTACT_ELLIPTIC) ? d->con_dim[d->efc_id[i]] : 1;
      double res[6], A[36], old[6];
      for (int j = 0; j < dim; j++) {
        int r = i + j;
        res[j] = row_dot(d, r, qacc) - d->efc_aref[r] + d->efc_R[r] * d->efc_force[r];
        for (int k = 0; k < dim; k++) A[j * dim + k] = d->efc_A[3 * r + k];
        old[j] = d->efc_force[r];
        flops += 2.0 * d->efc_rownnz[r] + 3;
      }
      improvement -= pgs_block(d, i, dim, res, A);
      for (int j = 0; j < dim; j++) {
        int r = i + j; double dl = d->efc_force[r] - old[j];
        if (dl != 0) { int adr = d->efc_rowadr[r]; for (int q = 0; q < d->efc_rownnz[r]; q++) qacc[d->efc_colind[adr + q]] += d->efc_B[adr + q] * dl; flops += 2.0 * d->efc_rownnz[r]; }
      }
      flops += dim == 1 ? 6 : 120;
      i += dim - 1;
    }
    improvement *= scale;
    iter++;
    if (improvement < m->tolerance) break;
  }
  d->solver_iter = iter; d->flops += flops;
}

/* literal mj_solPGS on an explicit efc_AR = J M^-1 J^T + diag(R) and efc_b (validation mode only) */
static void solve_pgs_dense(sgo_world* d) {
  const sgo_model* m = d->m; int nefc = d->nefc, nv = m->nv;
  double* AR = (double*)calloc((size_t)nefc * nefc, sizeof(double));
  double* Jd = (double*)calloc((size_t)nefc * nv, sizeof(double));
  double* Bd = (double*)calloc((size_t)nefc * nv, sizeof(double));
  for (int r = 0; r < nefc; r++) for (int q = 0; q < d->efc_rownnz[r]; q++) { int c = d->efc_colind[d->efc_rowadr[r] + q]; Jd[(size_t)r * nv + c] = d->efc_J[d->efc_rowadr[r] + q]; Bd[(size_t)r * nv + c] = d->efc_B[d->efc_rowadr[r] + q]; }
  for (int r = 0; r < nefc; r++) for (int c = 0; c < nefc; c++) { double s = 0; for (int k = 0; k < nv; k++) s += Jd[(size_t)r * nv + k] * Bd[(size_t)c * nv + k]; AR[(size_t)r * nefc + c] = s + (r == c ? d->efc_R[r] : 0); }
  double scale = 1 / (m->meaninertia * (nv > 1 ? nv : 1));
  int iter = 0;
  while (iter < m->iterations) {
    double improvement = 0;
    for (int i = 0; i < nefc; i++) {
      int dim = (d->efc_type[i] == CNSTR_CONTACT_ELLIPTIC) ? d->con_dim[d->efc_id[i]] : 1;
      double res[6], A[36];
      for (int j = 0; j < dim; j++) {
        double s = d->efc_b[i + j];
        for (int k = 0; k < nefc; k++) s += AR[(size_t)(i + j) * nefc + k] * d->efc_force[k];
        res[j] = s;
        for (int k = 0; k < dim; k++) A[j * dim + k] = AR[(size_t)(i + j) * nefc + i + k];
      }
      improvement -= pgs_block(d, i, dim, res, A);
      i += dim - 1;
    }
    improvement *= scale; iter++;
    if (improvement < m->tolerance) break;
  }
  d->solver_iter = iter;
  free(AR); free(Jd); free(Bd);
}

/* mj_fwdConstraint: b, warm start, PGS, dualFinish */
static void fwd_constraint(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv, nefc = d->nefc;
  /* mj_fwdConstraint saves the result for the next step's warm start itself (both exits), so mj_forward alone -- ManEnv.reset's
   * sim.forward(), ref: manenv.py:57-58 -- already seeds the first step; the integrator does not touch qacc_warmstart */
  if (!nefc) {
    memcpy(d->qacc, d->qacc_smooth, sizeof(double) * nv); memcpy(d->qacc_warmstart, d->qacc_smooth, sizeof(double) * nv);
    memset(d->qfrc_constraint, 0, sizeof(double) * nv); d->solver_iter = 0; return;
  }
  for (int i = 0; i < nefc; i++) d->efc_b[i] = row_dot(d, i, d->qacc_smooth) - d->efc_aref[i];
  /* warm start: forces from qacc_warmstart, kept only if the dual cost is not positive */
  double* jar = (double*)malloc(sizeof(double) * nefc);
  for (int i = 0; i < nefc; i++) jar[i] = row_dot(d, i, d->qacc_warmstart) - d->efc_aref[i];
  constraint_update(d, jar);
  free(jar);
  /* cost = f.b + 0.5 f'(J M^-1 J')f + 0.5 sum R f^2, with w = M^-1 J' f */
  double* w = d->scratch; memset(w, 0, sizeof(double) * nv);
  double cost = 0;
  for (int i = 0; i < nefc; i++) {
    double f = d->efc_force[i]; cost += f * d->efc_b[i] + 0.5 * d->efc_R[i] * f * f;
    if (f != 0) { int adr = d->efc_rowadr[i]; for (int q = 0; q < d->efc_rownnz[i]; q++) w[d->efc_colind[adr + q]] += d->efc_B[adr + q] * f; }
  }
  for (int i = 0; i < nefc; i++) cost += 0.5 * d->efc_force[i] * row_dot(d, i, w);
  if (cost > 0) { memset(d->efc_force, 0, sizeof(double) * nefc); memset(w, 0, sizeof(double) * nv); }
  for (int i = 0; i < nv; i++) d->qacc[i] = d->qacc_smooth[i] + w[i];
  if (d->dense) solve_pgs_dense(d); else solve_pgs(d);
  /* dualFinish: qfrc_constraint = J' f ; qacc = qacc_smooth + M^-1 qfrc_constraint */
  memset(d->qfrc_constraint, 0, sizeof(double) * nv);
  memset(w, 0, sizeof(double) * nv);
  for (int i = 0; i < nefc; i++) {
    double f = d->efc_force[i]; if (f == 0) continue;
    int adr = d->efc_rowadr[i];
    for (int q = 0; q < d->efc_rownnz[i]; q++) { int c = d->efc_colind[adr + q]; d->qfrc_constraint[c] += d->efc_J[adr + q] * f; w[c] += d->efc_B[adr + q] * f; }
  }
  for (int i = 0; i < nv; i++) d->qacc[i] = d->qacc_smooth[i] + w[i];
  memcpy(d->qacc_warmstart, d->qacc, sizeof(double) * nv);
}

/* ------------------------------------------------------------------------------------------ */
/* mj_forward / mj_step                                                                        */
/* ------------------------------------------------------------------------------------------ */
static int bad(const double* x, int n) { for (int i = 0; i < n; i++) if (!(fabs(x[i]) <= MAXVAL)) return 1; return 0; }

void sgo_forward(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv;
  d->flops = 0;
  /* fwdPosition */
  kinematics(d); com_pos(d); tendon(d); crb_factor(d); collision(d); make_constraint(d);
  /* fwdVelocity */
  for (int t = 0; t < m->ntendon; t++) { double s = 0; for (int i = 0; i < nv; i++) s += d->ten_J[(size_t)t * nv + i] * d->qvel[i]; d->ten_velocity[t] = s; }
  com_vel(d); passive(d);
  make_impedance(d);
  for (int i = 0; i < d->nefc; i++) { d->efc_vel[i] = row_dot(d, i, d->qvel); d->efc_aref[i] = -d->efc_Bd[i] * d->efc_vel[i] - d->efc_K[i] * d->efc_imp[i] * (d->efc_pos[i] - d->efc_margin[i]); }
  rne(d, 0, d->qfrc_bias);
  sensors(d, 0);
  /* fwdActuation (cylinder = filter dynamics, fixed gain, affine bias) */
  memset(d->qfrc_actuator, 0, sizeof(double) * nv);
  for (int u = 0; u < m->nu; u++) {
    int t = m->actuator_trnid[u]; double gear = m->actuator_gear[u];
    d->act_dot[u] = (d->ctrl[u] - d->act[u]) / fmax(MINVAL, m->actuator_timeconst[u]);
    double len = gear * d->ten_length[t], vel = gear * d->ten_velocity[t];
    double f = m->actuator_gain[u] * d->act[u] + m->actuator_bias[3 * u] + m->actuator_bias[3 * u + 1] * len + m->actuator_bias[3 * u + 2] * vel;
    d->actuator_force[u] = f;
    for (int i = 0; i < nv; i++) d->qfrc_actuator[i] += gear * d->ten_J[(size_t)t * nv + i] * f;
  }
  /* fwdAcceleration */
  for (int i = 0; i < nv; i++) { d->qfrc_smooth[i] = d->qfrc_passive[i] - d->qfrc_bias[i] + d->qfrc_actuator[i]; d->qacc_smooth[i] = d->qfrc_smooth[i]; }
  solve_ld(d, d->qacc_smooth);
  fwd_constraint(d);
  /* sensorAcc */
  rne(d, 1, NULL);
  sensors(d, 1);
  d->steps_total++; d->flops_pgs_total += d->flops;
  d->flops_total += d->flops + 60.0 * nv + 20.0 * m->neq + 400.0 * d->ncand + 100.0 * d->ncon + 3000.0;
}

/* mj_Euler with implicit joint damping */
static void euler(sgo_world* d) {
  const sgo_model* m = d->m; int nv = m->nv; double h = m->timestep;
  double* qacc = d->scratch;
  int damped = 0; for (int i = 0; i < nv; i++) if (d->dof_damping[i] > 0) { damped = 1; break; }
  if (!damped) memcpy(qacc, d->qacc, sizeof(double) * nv);
  else {
    double* MhB = (double*)malloc(sizeof(double) * m->nM * 2 + sizeof(double) * nv);
    double* LD = MhB + m->nM; double* dinv = LD + m->nM;
    memcpy(MhB, d->qM, sizeof(double) * m->nM);
    for (int i = 0; i < nv; i++) MhB[m->dof_Madr[i]] += h * d->dof_damping[i];
    factor_i(m, MhB, LD, dinv);
    for (int i = 0; i < nv; i++) qacc[i] = d->qfrc_smooth[i] + d->qfrc_constraint[i];
    /* solve with the temporary factor */
    double *sLD = d->qLD, *sDi = d->qLDiagInv;
    d->qLD = LD; d->qLDiagInv = dinv; solve_ld(d, qacc);
    if (d->implicit_tendon) {
      /* HYPOTHESIS SWITCH (SURVEY App. E candidate, off by default, NOT what MuJoCo's Euler integrator is recalled to do):
       * tendon dampers treated like joint dampers, (M + h diag(d) + h sum_t b_t J_t' J_t) qacc' = rhs, one rank-1
       * Sherman-Morrison update per damped tendon on top of the solve above */
      double* y = (double*)malloc(sizeof(double) * nv);
      for (int t = 0; t < m->ntendon; t++) {
        double b = d->tendon_damping[t];
        if (!(b > 0)) continue;
        const double* J = d->ten_J + (size_t)t * nv;
        memcpy(y, J, sizeof(double) * nv);
        solve_ld(d, y);
        double jx = 0, jy = 0;
        for (int i = 0; i < nv; i++) { jx += J[i] * qacc[i]; jy += J[i] * y[i]; }
        double c = h * b * jx / (1.0 + h * b * jy);
        for (int i = 0; i < nv; i++) qacc[i] -= c * y[i];
      }
      free(y);
    }
    d->qLD = sLD; d->qLDiagInv = sDi;
    free(MhB);
  }
  for (int u = 0; u < m->nu; u++) d->act[u] += h * d->act_dot[u];
  for (int i = 0; i < nv; i++) { d->qvel[i] += h * qacc[i]; d->qpos[i] += h * d->qvel[i]; }
  d->time += h;
}

int sgo_step(sgo_world* d) {
  const sgo_model* m = d->m;
  d->step_status = 0;
  if (bad(d->qpos, m->nv) || bad(d->qvel, m->nv)) { d->step_status |= SGO_ST_DIVERGED; sgo_reset(d); }
  sgo_forward(d);
  if (bad(d->qacc, m->nv)) { d->step_status |= SGO_ST_DIVERGED; sgo_reset(d); sgo_forward(d); }
  euler(d);
  d->status |= d->step_status;
  return d->step_status;
}

/* ------------------------------------------------------------------------------------------ */
/* accessors                                                                                   */
/* ------------------------------------------------------------------------------------------ */
void sgo_get_state(const sgo_world* d, double* qpos, double* qvel, double* act, double* ws) {
  int nv = d->m->nv;
  if (qpos) memcpy(qpos, d->qpos, sizeof(double) * nv);
  if (qvel) memcpy(qvel, d->qvel, sizeof(double) * nv);
  if (act) memcpy(act, d->act, sizeof(double) * d->m->nu);
  if (ws) memcpy(ws, d->qacc_warmstart, sizeof(double) * nv);
}
void sgo_set_state(sgo_world* d, const double* qpos, const double* qvel, const double* act, const double* ws) {
  int nv = d->m->nv;
  if (qpos) memcpy(d->qpos, qpos, sizeof(double) * nv);
  if (qvel) memcpy(d->qvel, qvel, sizeof(double) * nv);
  if (act) memcpy(d->act, act, sizeof(double) * d->m->nu);
  if (ws) memcpy(d->qacc_warmstart, ws, sizeof(double) * nv);
}
void sgo_get_sensordata(const sgo_world* d, double* out) { memcpy(out, d->sensordata, sizeof(double) * d->m->nsensordata); }
int sgo_get_touch_mask(const sgo_world* d) { return d->touch_mask; }
int sgo_get_int(const sgo_world* d, const char* k) {
  if (!strcmp(k, "ncon")) return d->ncon;
  if (!strcmp(k, "nefc")) return d->nefc;
  if (!strcmp(k, "ne")) return d->ne;
  if (!strcmp(k, "nl")) return d->nl;
  if (!strcmp(k, "solver_iter")) return d->solver_iter;
  return -1;
}
static int copy_out(double* out, int cap, const double* src, int n) { if (out) memcpy(out, src, sizeof(double) * (n < cap ? n : cap)); return n; }
static int copy_outi(double* out, int cap, const int* src, int n) { if (out) for (int i = 0; i < n && i < cap; i++) out[i] = src[i]; return n; }
int sgo_get_array(const sgo_world* d, const char* k, double* out, int cap) {
  const sgo_model* m = d->m; int nv = m->nv;
#define A(name, ptr, n) if (!strcmp(k, name)) return copy_out(out, cap, ptr, n)
  A("qpos", d->qpos, nv); A("qvel", d->qvel, nv); A("qacc", d->qacc, nv); A("qacc_smooth", d->qacc_smooth, nv);
  A("qacc_warmstart", d->qacc_warmstart, nv); A("act", d->act, m->nu); A("act_dot", d->act_dot, m->nu);
  A("qfrc_bias", d->qfrc_bias, nv); A("qfrc_passive", d->qfrc_passive, nv); A("qfrc_actuator", d->qfrc_actuator, nv);
  A("qfrc_smooth", d->qfrc_smooth, nv); A("qfrc_constraint", d->qfrc_constraint, nv); A("qM", d->qM, m->nM);
  A("xpos", d->xpos, 3 * m->nbody); A("xmat", d->xmat, 9 * m->nbody); A("xipos", d->xipos, 3 * m->nbody);
  A("geom_xpos", d->geom_xpos, 3 * m->ngeom); A("geom_xmat", d->geom_xmat, 9 * m->ngeom);
  A("site_xpos", d->site_xpos, 3 * m->nsite); A("site_xmat", d->site_xmat, 9 * m->nsite);
  A("subtree_com", d->subtree_com, 3 * m->nbody); A("cdof", d->cdof, 6 * nv); A("cvel", d->cvel, 6 * m->nbody);
  A("cacc", d->cacc, 6 * m->nbody);
  A("ten_length", d->ten_length, m->ntendon); A("ten_velocity", d->ten_velocity, m->ntendon); A("ten_J", d->ten_J, m->ntendon * nv);
  A("efc_pos", d->efc_pos, d->nefc); A("efc_aref", d->efc_aref, d->nefc); A("efc_R", d->efc_R, d->nefc);
  A("efc_force", d->efc_force, d->nefc); A("efc_b", d->efc_b, d->nefc); A("efc_imp", d->efc_imp, d->nefc);
  A("efc_K", d->efc_K, d->nefc); A("efc_B", d->efc_Bd, d->nefc); A("efc_diagApprox", d->efc_diagApprox, d->nefc);
  A("efc_vel", d->efc_vel, d->nefc);
  A("con_dist", d->con_dist, d->ncon); A("con_pos", d->con_pos, 3 * d->ncon); A("con_frame", d->con_frame, 9 * d->ncon);
  A("con_mu", d->con_mu, d->ncon); A("sensordata", d->sensordata, m->nsensordata); A("tree_Minv", d->tree_Minv, m->ntreeminv);
#undef A
  if (!strcmp(k, "efc_type")) return copy_outi(out, cap, d->efc_type, d->nefc);
  if (!strcmp(k, "efc_id")) return copy_outi(out, cap, d->efc_id, d->nefc);
  if (!strcmp(k, "con_geom1")) return copy_outi(out, cap, d->con_geom1, d->ncon);
  if (!strcmp(k, "con_geom2")) return copy_outi(out, cap, d->con_geom2, d->ncon);
  if (!strcmp(k, "con_exclude")) return copy_outi(out, cap, d->con_exclude, d->ncon);
  if (!strcmp(k, "pair_g1")) return copy_outi(out, cap, m->pair_g1, m->npair);
  if (!strcmp(k, "pair_g2")) return copy_outi(out, cap, m->pair_g2, m->npair);
  if (!strcmp(k, "efc_J")) {   /* dense nefc x nv */
    int n = d->nefc * nv;
    if (out) { for (int i = 0; i < n && i < cap; i++) out[i] = 0; for (int r = 0; r < d->nefc; r++) for (int q = 0; q < d->efc_rownnz[r]; q++) { size_t idx = (size_t)r * nv + d->efc_colind[d->efc_rowadr[r] + q]; if ((int)idx < cap) out[idx] = d->efc_J[d->efc_rowadr[r] + q]; } }
    return n;
  }
  if (!strcmp(k, "M_dense")) {
    int n = nv * nv;
    if (out && cap >= n) { memset(out, 0, sizeof(double) * n); for (int i = 0; i < nv; i++) { int adr = m->dof_Madr[i], j = i; while (j >= 0) { out[i * nv + j] = out[j * nv + i] = d->qM[adr++]; j = m->dof_parentid[j]; } } }
    return n;
  }
  return -1;
}

/* ------------------------------------------------------------------------------------------ */
/* episode protocol of create_dataset.log_into_file (ref: create_dataset.py:33-60)              */
/* ------------------------------------------------------------------------------------------ */
static int env_step(sgo_world* d, int nsub, double* row, int* touch) {
  int st = 0;
  for (int s = 0; s < nsub; s++) st |= sgo_step(d);
  if (row) memcpy(row, d->sensordata, sizeof(double) * d->m->nsensordata);
  if (touch) *touch = d->touch_mask;
  return st;
}
int sgo_episode(sgo_world* d, int sim_start, int sim_step, int n_settle, int n_iter, int open_close_div,
                double ctrl_mag, double* out, int* touch) {
  const sgo_model* m = d->m; int ns = m->nsensordata, st = 0, row = 0, closing = 1;
  sgo_reset(d);                       /* ManEnv.reset: sim.reset(); sim.forward(); step(sim_start) */
  sgo_forward(d);
  if (sim_start > 0) st |= env_step(d, sim_start, NULL, NULL);
  for (int i = 0; i < n_settle; i++, row++) st |= env_step(d, sim_step, out + (size_t)row * ns, touch ? touch + row : NULL);
  for (int u = 0; u < m->nu; u++) d->ctrl[u] = -ctrl_mag;      /* close_hand */
  closing = 1;
  for (int i = 0; i < n_iter; i++, row++) {
    if (open_close_div > 0 && i % open_close_div == 0 && i > 0) {   /* toggle_grip */
      closing = !closing;
      for (int u = 0; u < m->nu; u++) d->ctrl[u] = closing ? -ctrl_mag : ctrl_mag;
    }
    st |= env_step(d, sim_step, out + (size_t)row * ns, touch ? touch + row : NULL);
  }
  return st;
}

/* ------------------------------------------------------------------------------------------ */
/* test hooks (known-answer tests of the static helpers)                                       */
/* ------------------------------------------------------------------------------------------ */
/* out: n, then per contact dist, pos[3], normal[3] */
int sgo_test_capsule_box(const double* cpos, const double* cmat, const double* csize, const double* bpos,
                         const double* bmat, const double* bsize, double* out) {
  rawcon rc[2];
  int n = capsule_box(rc, 0.0, cpos, cmat, csize, bpos, bmat, bsize);
  for (int i = 0; i < n; i++) { out[7 * i] = rc[i].dist; memcpy(out + 7 * i + 1, rc[i].pos, 3 * sizeof(double)); memcpy(out + 7 * i + 4, rc[i].frame, 3 * sizeof(double)); }
  return n;
}
int sgo_test_sphere_box(const double* spos, double radius, const double* bpos, const double* bmat, const double* bsize, double* out) {
  rawcon rc;
  int n = sphere_box(&rc, 0.0, spos, radius, bpos, bmat, bsize);
  if (n) { out[0] = rc.dist; memcpy(out + 1, rc.pos, 3 * sizeof(double)); memcpy(out + 4, rc.frame, 3 * sizeof(double)); }
  return n;
}
int sgo_test_qcqp2(const double* A, const double* b, const double* d, double r, double* res) { return qcqp2(res, A, b, d, r); }
double sgo_test_impedance(const double* solimp, double pos, double margin) { return impedance(solimp, pos, margin); }
int sgo_test_box_box(const double* p1, const double* R1, const double* s1, const double* p2, const double* R2, const double* s2) { return box_box_overlap(p1, R1, s1, p2, R2, s2); }
void sgo_test_make_frame(double* frame) { make_frame(frame); }
