/* sg_oracle.h -- TEST INFRASTRUCTURE ONLY (see header of sg_oracle.c).
 *
 * fp64 single-world CPU restatement of the MuJoCo step loop the reference drives through
 * mujoco-py (ref: environment/manenv.py:44-63).  PARITY UNPINNED: the reference ships no golden
 * vectors and MuJoCo itself is not installable here.
 */
#ifndef SG_ORACLE_H
#define SG_ORACLE_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sgo_model sgo_model;
typedef struct sgo_world sgo_world;

/* status bits returned by sgo_step / accumulated in sgo_status */
#define SGO_ST_DIVERGED   1   /* NaN / |x|>1e10 in qpos,qvel,qacc -> data was reset (mj_checkPos/Vel/Acc) */
#define SGO_ST_CON_FULL   2   /* more contacts than nconmax: extra contacts dropped            */
#define SGO_ST_EFC_FULL   4   /* more constraint rows than njmax                               */
#define SGO_ST_BOXBOX     8   /* a box-box pair overlaps: not restated (tested-and-rejected only) */

sgo_model* sgo_model_load(const void* blob, size_t nbytes);
void       sgo_model_free(sgo_model* m);
int        sgo_model_int(const sgo_model* m, const char* key);   /* nv, nbody, ngeom, neq, nu, nsensordata, ntendon, njnt, npair */

sgo_world* sgo_world_create(const sgo_model* m);
void       sgo_world_free(sgo_world* w);

/* per-world model parameters (what ManEnv.set_new_stiffness writes, ref: manenv.py:103-109) */
void sgo_set_jnt_stiffness(sgo_world* w, int jnt, double k);
void sgo_set_tendon_stiffness(sgo_world* w, int tendon, double k);
void sgo_set_dof_damping(sgo_world* w, int dof, double d);
void sgo_set_tendon_damping(sgo_world* w, int tendon, double d);
void sgo_set_body_pos(sgo_world* w, int body, const double* xyz);
void sgo_set_ctrl(sgo_world* w, const double* ctrl);
void sgo_set_capsule_box_single(sgo_world* w, int on); /* sensitivity study only: drop the second capsule-box contact */
void sgo_set_implicit_tendon_damping(sgo_world* w, int on); /* hypothesis switch, off by default */
void sgo_set_dense_solver(sgo_world* w, int on);   /* literal efc_AR PGS (slow; validation of the matrix-free form) */
void sgo_set_geom_mask(sgo_world* w, const int* mask); /* per-geom name bitmask for the contact flag */

void sgo_reset(sgo_world* w);      /* mj_resetData */
void sgo_forward(sgo_world* w);    /* mj_forward   */
int  sgo_step(sgo_world* w);       /* mj_step; returns status bits of this step */
int  sgo_status(const sgo_world* w);

/* state access (copies) */
void sgo_get_state(const sgo_world* w, double* qpos, double* qvel, double* act, double* qacc_warmstart);
void sgo_set_state(sgo_world* w, const double* qpos, const double* qvel, const double* act, const double* qacc_warmstart);
void sgo_get_sensordata(const sgo_world* w, double* out);
int  sgo_get_touch_mask(const sgo_world* w);       /* OR over contacts involving an object geom of (mask>>1) */
int  sgo_get_int(const sgo_world* w, const char* key);          /* ncon, nefc, ne, nl, solver_iter */
int  sgo_get_array(const sgo_world* w, const char* key, double* out, int cap); /* returns count */

/* whole squeeze episode of create_dataset.log_into_file (ref: create_dataset.py:33-60).
 * out: nrows*nsensordata doubles; touch: nrows ints.  Returns accumulated status bits. */
int sgo_episode(sgo_world* w, int sim_start, int sim_step, int n_settle, int n_iter,
                int open_close_div, double ctrl_mag, double* out, int* touch);

/* op counter (algorithmic flop estimate of the last step, counted in the PGS and its setup) */
double sgo_last_step_flops(const sgo_world* w);
/* accumulated since the last reset: out[0] forwards, out[1] PGS flops, out[2] all-stage flops (see sg_oracle.c) */
void sgo_flops_get(const sgo_world* w, double* out);
void sgo_flops_reset(sgo_world* w);

#ifdef __cplusplus
}
#endif
#endif
