// Flat device model + env configuration consumed by the fused step kernel (engine.cuh).
// POD only: uploaded once per batch with cudaMemcpy, read through the read-only cache.
#pragma once
#include <cstdint>

namespace ur3e {

// Static capacities of the device tables.  They are sized tightly for the reference scenes (after fixed-body merging): the tables are
// read through the ~24 KB of L1 that the shared-memory carve-out leaves, so every unused slot costs cache.
#ifndef UR3E_MAXB
#define UR3E_MAXB 20
#define UR3E_MAXG 14
#define UR3E_MAXPAIR 48
#define UR3E_MAXNM 104
#endif
constexpr int MAXB = UR3E_MAXB;     // bodies incl. world (main.xml merges to 19)
constexpr int MAXV = 20;     // dofs
constexpr int MAXQ = 21;     // generalized positions
constexpr int MAXU = 7;      // actuators
constexpr int MAXG = UR3E_MAXG;     // collidable primitive geoms (main.xml: 14)
constexpr int MAXPAIR = UR3E_MAXPAIR;  // candidate geom pairs (main.xml: 48)
constexpr int MAXEQ = 3;     // equality constraints
constexpr int MAXSITE = 4;   // tracked sites (tcp, handle, pad)
constexpr int MAXNM = UR3E_MAXNM;   // nnz of lower-triangular M (main.xml: 102)
constexpr int MAXKEY = 2;
constexpr int MAXTQ = 6;     // torque sensors (main.xml: one per arm joint)
constexpr int MAXANC = 12;    // ancestors of a dof in the dof tree
constexpr int MAXCON = 32;   // contacts per env
constexpr int MAXEFC = 112;  // constraint rows per env

enum JointKind { JK_NONE = 0, JK_HINGE = 1, JK_FREE = 2 };
enum GeomKind { GK_PLANE = 0, GK_BOX = 1 };
enum EqKind { EK_CONNECT = 0, EK_JOINT = 1 };
// which of the bodies the reference's contact predicates name a collidable geom belongs to (utils/gym_utils.py:108-201)
enum GeomFlag { GF_ARM = 1 /* robot_base subtree */, GF_GRIPPER = 2 /* robotiq_base_mount subtree */, GF_TABLE = 4, GF_MUG = 8, GF_LPAD = 16, GF_RPAD = 32 };

// controller evaluated inside the kernel (reference controller/controller_func.py)
enum CtrlMode {
  CTRL_RAW = 0,       // action = actuator ctrl vector (imitation_env_direct.py:90)
  CTRL_PD_JOINT = 1,  // pd_joint_ctrl + move_j.get_joint_delta (controller_func.py:128-167, move_j.py:14-38)
  CTRL_PID_TASK = 2,  // pid_task_ctrl, action = full 7-vector trajectory point (move_l_task.py:55-69)
  CTRL_PID_TASK_ENV = 3,  // pid_task_ctrl, action = [x,y,z,grip], fixed tool orientation (ur3e_env2.py:72-82)
  CTRL_PINV = 4       // move_l.ctrl: pinv(J) IK + two joint PDs (move_l.py:15-78)
};
enum ObsKind { OBS_STATE = 0 /* qpos,qvel */, OBS_V2 = 1 /* 24 */, OBS_V0 = 2 /* 13 */, OBS_DIRECT = 3 /* 13 */ };
enum RewardKind { REW_NONE = 0, REW_V2 = 1, REW_V0 = 2, REW_MINUS1 = 3 };
enum TermKind { TERM_NONE = 0, TERM_V2 = 1, TERM_V0 = 2 };
enum ResetNoise { NOISE_NONE = 0, NOISE_LOW = 1, NOISE_MED = 2, NOISE_HIGH = 3 };

template <typename Real>
struct DevModel {
  // sizes
  int nq, nv, nu, nbody, nlevel, ngeom, npair, neq, nsite, nM, nkey, nfl;
  int has_damping, pad_;
  Real timestep, gravity[3], impratio, meaninertia;
  // bodies (index 0 = world)
  int body_parent[MAXB], body_level[MAXB], body_jkind[MAXB], body_qadr[MAXB], body_dadr[MAXB], body_lastdof[MAXB], body_root[MAXB];
  uint32_t body_dofmask[MAXB];  // bit d set when dof d moves the body
  int lev_start[MAXB + 1], lev_body[MAXB];   // bodies grouped by tree depth: level l (>= 1) = lev_body[lev_start[l] .. lev_start[l+1])
  // constant rotations are stored as row-major 3x3 matrices (body frame in the parent, inertial frame in the body)
  Real body_pos[MAXB][3], body_mat[MAXB][9], body_ipos[MAXB][3], body_imat[MAXB][9], body_mass[MAXB], body_inertia[MAXB][3];
  Real body_invw[MAXB][2];
  Real jnt_pos[MAXB][3], jnt_axis[MAXB][3], jnt_q0[MAXB];  // the (single) joint of each body
  // dofs
  int dof_body[MAXV], dof_parent[MAXV], dof_qadr[MAXV], dof_limited[MAXV], dof_free_k[MAXV];  // free_k: -1 hinge, 0..5 component of a free joint
  Real dof_armature[MAXV], dof_damping[MAXV], dof_frictionloss[MAXV], dof_invw[MAXV], dof_stiffness[MAXV], dof_springref[MAXV];
  Real dof_range[MAXV][2], dof_margin[MAXV], dof_lim_solref[MAXV][2], dof_lim_solimp[MAXV][5], dof_fl_solref[MAXV][2], dof_fl_solimp[MAXV][5];
  int dof_nanc[MAXV], dof_anc[MAXV][MAXANC], max_nanc;   // dof-tree ancestors, nearest first (sparse L^T D L of the Euler solve)
  int fl_dof[MAXV];  // dofs with frictionloss, in order
  int dof_flrow[MAXV];  // position of the dof in fl_dof, -1 when it has no friction loss
  // lower-triangular sparsity of M: (i, j) with j ancestor-or-self of i
  int M_i[MAXNM], M_j[MAXNM];
  int tri_ab[(MAXV + 1) * (MAXV + 2) / 2];  // (a << 8 | b), b <= a, row-major lower triangle of the (nv+1)^2 augmented matrix
  // collidable geoms
  int geom_body[MAXG], geom_kind[MAXG], geom_src[MAXG];   // geom_src: index of the geom in the loaded model
  int geom_flags[MAXG];   // GeomFlag bits by the geom's body in the loaded (unmerged) model; set at batch creation
  Real geom_pos[MAXG][3], geom_mat[MAXG][9], geom_size[MAXG][3], geom_rbound[MAXG];
  // candidate pairs (geom1 = plane for plane-box)
  int pair_g1[MAXPAIR], pair_g2[MAXPAIR], pair_src_g1[MAXPAIR], pair_src_g2[MAXPAIR];
  int pair_code[MAXPAIR];     // broad phase: g1 | g2 << 8 | (g1 is a plane) << 16
  Real pair_rsum[MAXPAIR];    // broad phase: bounding radius of g1 (0 for a plane) + bounding radius of g2 + margin
  Real geom_nrm[MAXG][3];     // planes (always on static bodies): world normal
  Real pair_friction[MAXPAIR][2], pair_solref[MAXPAIR][2], pair_solimp[MAXPAIR][5], pair_margin[MAXPAIR], pair_includemargin[MAXPAIR], pair_invw[MAXPAIR];
  // equality
  int eq_kind[MAXEQ], eq_o1[MAXEQ], eq_o2[MAXEQ];
  Real eq_data[MAXEQ][6], eq_solref[MAXEQ][2], eq_solimp[MAXEQ][5], eq_invw[MAXEQ];
  // tracked sites
  int site_body[MAXSITE];
  Real site_pos[MAXSITE][3], site_mat[MAXSITE][9], site_size[MAXSITE][3];
  // torque sensors (logging only): body, and the site frame in that body
  int ntq, tq_body[MAXTQ];
  Real tq_pos[MAXTQ][3], tq_mat[MAXTQ][9];
  // actuators: force = gain*ctrl + b0 + b1*len + b2*vel ; moment over at most two dofs
  int act_dof[MAXU][2], act_ctrllimited[MAXU], act_forcelimited[MAXU];
  int dof_nact[MAXV], dof_act[MAXV][2]; Real dof_actcoef[MAXV][2];   // transposed transmission: the (at most two) actuators acting on each dof
  int split;       // first dof of the last kinematic tree when the model has more than one (else nv): M is block diagonal across it
  int ndeq, nej;   // rows of the connect equalities (3 each) / number of joint equalities
  Real act_coef[MAXU][2], act_gain[MAXU], act_bias[MAXU][3], act_ctrlrange[MAXU][2], act_forcerange[MAXU][2];
  // reset
  Real qpos0[MAXQ], key_qpos[MAXKEY][MAXQ], key_qvel[MAXKEY][MAXV];
};

template <typename Real>
struct EnvCfg {
  int ctrl_mode, obs_kind, reward_kind, term_kind;
  int frame_skip, act_dim, obs_dim, max_steps;
  int trunc_after_increment;  // v2 increments t before the truncation test (ur3e_env2.py:89-92); others test first
  int reset_key, reset_noise, auto_reset;
  int site_tcp, site_mug, site_pad, body_mug, body_ghost, body_lpad, body_rpad, finger_q;  // model indices (-1 when absent); sites index the tracked-site list
  int body_table, pad0_, pad1_, pad_;
  uint32_t gripper_mask;   // bit b: (merged) body b belongs to the robotiq_base_mount subtree (gym_utils.py:133-143)
  Real gains[24];          // CTRL_PID_TASK*: kp_pos[3] kd_pos[3] kp_rot[3] kd_rot[3]; CTRL_PD_JOINT: kp[6] kd[6]; CTRL_PINV: kp_p[6] kd_p[6] kp_r[6] kd_r[6]
  Real tool_rotvec[3];     // ur3e_env2.py:74
  Real act_low[8], act_high[8];
  Real mug_size[3];
  Real ghost_pos[3];
  Real topple_z;           // max(size_x, size_y) of the mug box (gym_utils.py:8-17)
};

constexpr int NSENSOR = 46;  // logging record: 7 x actuatorfrc, touch right_pad1_contact, touch left_pad1_contact, tcp xpos (3), tcp xmat (9), d.ctrl (7), 6 x torque (3)
constexpr int CACHE_SIZE = 3 + 9 + 36 + 6;  // tcp_pos, tcp_mat, J_arm (6x6: rows px,py,pz,rx,ry,rz), qfrc_bias[:6]

}  // namespace ur3e
