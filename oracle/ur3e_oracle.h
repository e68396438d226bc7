/* ORACLE (test infrastructure, NOT product code).
 *
 * float64 CPU restatement of the reference's hot path
 *   controller -> mj_step x frame_skip -> obs/reward/done
 * (reference gymnasium_env/envs/ur3e_env2.py:72-99, controller/controller_func.py:68-117,
 * 128-167, and the mujoco==3.3.3 mj_step it calls, restated from SURVEY App. B).
 *
 * Parity status: UNPINNED against a real MuJoCo build (not installable here);
 * pinned by the tcp@'down' golden vector (reference assets/main.xml:415), analytic
 * invariants (tests/test_oracle_*.py) and the reference's own Python controller code run
 * on top of this library through oracle/mujoco_shim.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference legs may use it.
 */
#ifndef UR3E_ORACLE_H
#define UR3E_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define O_MAXCON 64
#define O_MAXEFC 256

/* model arrays (MuJoCo names). X(name) */
#define O_MODEL_INT_FIELDS(X) \
  X(body_parentid) X(body_rootid) X(body_weldid) X(body_jntadr) X(body_jntnum) X(body_dofadr) X(body_dofnum) \
  X(jnt_type) X(jnt_bodyid) X(jnt_qposadr) X(jnt_dofadr) X(jnt_limited) \
  X(dof_bodyid) X(dof_jntid) X(dof_parentid) \
  X(geom_type) X(geom_bodyid) X(site_bodyid) \
  X(tendon_adr) X(tendon_num) X(wrap_jnt) \
  X(eq_type) X(eq_obj1id) X(eq_obj2id) \
  X(actuator_trntype) X(actuator_trnid) X(actuator_ctrllimited) X(actuator_forcelimited) \
  X(pair_geom1) X(pair_geom2) X(pair_condim)

#define O_MODEL_DBL_FIELDS(X) \
  X(body_pos) X(body_quat) X(body_ipos) X(body_iquat) X(body_mass) X(body_inertia) X(body_invweight0) \
  X(jnt_pos) X(jnt_axis) X(jnt_range) X(jnt_stiffness) X(jnt_margin) X(jnt_solref) X(jnt_solimp) \
  X(qpos0) X(qpos_spring) \
  X(dof_armature) X(dof_damping) X(dof_frictionloss) X(dof_invweight0) X(dof_solref) X(dof_solimp) \
  X(geom_pos) X(geom_quat) X(geom_size) X(site_pos) X(site_quat) \
  X(wrap_coef) X(tendon_invweight0) \
  X(eq_data) X(eq_solref) X(eq_solimp) \
  X(actuator_gainprm) X(actuator_biasprm) X(actuator_ctrlrange) X(actuator_forcerange) X(actuator_gear) \
  X(pair_friction) X(pair_solref) X(pair_solimp) X(pair_margin) X(pair_gap) \
  X(key_qpos) X(key_qvel)

typedef struct OModel {
  int nq, nv, nu, nbody, njnt, ngeom, nsite, neq, ntendon, nwrap, npair, nkey;
  double timestep, gravity[3], impratio, tolerance, ls_tolerance, meaninertia;
  int cone_elliptic, iterations, ls_iterations;
#define X(n) int* n;
  O_MODEL_INT_FIELDS(X)
#undef X
#define X(n) double* n;
  O_MODEL_DBL_FIELDS(X)
#undef X
} OModel;

typedef struct OContact {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5], mu, includemargin;
  int geom1, geom2, dim, efc_address;
} OContact;

/* data arrays. X(name, count-expression in terms of m) */
#define O_DATA_DBL_FIELDS(X) \
  X(qpos, m->nq) X(qvel, m->nv) X(qacc, m->nv) X(qacc_warmstart, m->nv) X(ctrl, m->nu) \
  X(xpos, 3*m->nbody) X(xquat, 4*m->nbody) X(xmat, 9*m->nbody) X(xipos, 3*m->nbody) X(ximat, 9*m->nbody) \
  X(xanchor, 3*m->njnt) X(xaxis, 3*m->njnt) X(geom_xpos, 3*m->ngeom) X(geom_xmat, 9*m->ngeom) \
  X(site_xpos, 3*m->nsite) X(site_xmat, 9*m->nsite) X(subtree_com, 3*m->nbody) \
  X(cinert, 10*m->nbody) X(crb, 10*m->nbody) X(cdof, 6*m->nv) X(cdof_dot, 6*m->nv) \
  X(cvel, 6*m->nbody) X(cacc, 6*m->nbody) X(cfrc_body, 6*m->nbody) \
  X(qM, m->nv*m->nv) X(qL, m->nv*m->nv) X(qH, m->nv*m->nv) \
  X(ten_length, m->ntendon) X(ten_velocity, m->ntendon) X(ten_J, m->ntendon*m->nv) \
  X(actuator_length, m->nu) X(actuator_velocity, m->nu) X(actuator_force, m->nu) X(actuator_moment, m->nu*m->nv) \
  X(qfrc_passive, m->nv) X(qfrc_bias, m->nv) X(qfrc_actuator, m->nv) X(qfrc_smooth, m->nv) \
  X(qacc_smooth, m->nv) X(qfrc_constraint, m->nv) \
  X(efc_J, O_MAXEFC*m->nv) X(efc_pos, O_MAXEFC) X(efc_margin, O_MAXEFC) X(efc_frictionloss, O_MAXEFC) \
  X(efc_diagApprox, O_MAXEFC) X(efc_R, O_MAXEFC) X(efc_D, O_MAXEFC) X(efc_KBIP, 4*O_MAXEFC) \
  X(efc_vel, O_MAXEFC) X(efc_aref, O_MAXEFC) X(efc_force, O_MAXEFC)

#define O_DATA_INT_FIELDS(X) X(efc_type, O_MAXEFC) X(efc_id, O_MAXEFC) X(efc_state, O_MAXEFC)

typedef struct OData {
  double time;
  int ncon, nefc, solver_iter, warn_bad;   /* warn_bad: 1 pos, 2 vel, 4 acc (mj_check*) */
  int ne, nf, nl;                          /* #equality, #friction, #limit rows */
  double solver_cost;
#define X(n, c) double* n;
  O_DATA_DBL_FIELDS(X)
#undef X
#define X(n, c) int* n;
  O_DATA_INT_FIELDS(X)
#undef X
  OContact contact[O_MAXCON];
} OData;

enum { O_CNSTR_EQUALITY = 0, O_CNSTR_FRICTION_DOF = 1, O_CNSTR_LIMIT_JOINT = 3, O_CNSTR_CONTACT_ELLIPTIC = 7 };
enum { O_STATE_SATISFIED = 0, O_STATE_QUADRATIC = 1, O_STATE_LINEARNEG = 2, O_STATE_LINEARPOS = 3, O_STATE_CONE = 4 };

/* construction */
OModel* o_model_new(const int* sizes12, const double* opt /* timestep, g[3], impratio, tolerance, ls_tolerance, cone, iterations, ls_iterations */);
int o_model_set_int(OModel* m, const char* name, const int* src, int n);
int o_model_set_dbl(OModel* m, const char* name, const double* src, int n);
int o_model_get_dbl(OModel* m, const char* name, double** ptr);
void o_model_free(OModel* m);
void o_set_const(OModel* m);   /* mj_setConst: anchors, invweight0, meaninertia */
double o_model_meaninertia(const OModel* m);

OData* o_data_new(const OModel* m);
void o_data_free(OData* d);
int o_data_get_dbl(const OModel* m, OData* d, const char* name, double** ptr, int* n);
int o_data_get_int(const OModel* m, OData* d, const char* name, int** ptr, int* n);
OContact* o_data_contacts(OData* d);
int o_data_info(OData* d, int which); /* 0 ncon 1 nefc 2 solver_iter 3 warn_bad 4 ne 5 nf 6 nl */
double o_data_time(OData* d);

/* pipeline (mujoco.mj_* restated) */
void o_reset_data(const OModel* m, OData* d);
void o_forward(const OModel* m, OData* d);
void o_step(const OModel* m, OData* d);
void o_step_n(const OModel* m, OData* d, int nstep);
void o_jac(const OModel* m, const OData* d, double* jacp, double* jacr, const double* point, int body);
void o_jac_site(const OModel* m, const OData* d, double* jacp, double* jacr, int site);
void o_object_velocity_site(const OModel* m, const OData* d, int site, double* res6, int flg_local);
void o_fullM(const OModel* m, const OData* d, double* dst);

/* controllers (reference controller/controller_func.py) */
void o_pid_task_ctrl(const OModel* m, const OData* d, int tcp_site, const double* traj7, const double* gains12, double* u7);
void o_pd_joint_ctrl(const OModel* m, const OData* d, const double* target6, const double* kp6, const double* kd6, double* u6);
void o_rot_err(const double* xmat9, const double* rotvec_target3, double* err3);

#ifdef __cplusplus
}
#endif
#endif
