// rsrx_device.cuh — device-side model, shared-memory arena layout and small math
// for the fused Airbot env-step kernel (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/rsrx.h"

// The whole stepper is compiled twice into librsrx.so: as `rsrx` (RSRX_MAXC = 24 active contacts per env, the arena
// that fits 19 envs into one SM's shared memory) and as `rsrx_big` (rsrx_redo.cu: RSRX_MAXC = 4 * RSRX_MAXPAIR, i.e. every
// slot of every geom pair — it cannot overflow), which re-runs the rare env-steps the first one had to give up on.
#ifndef RSRX_NS
#define RSRX_NS rsrx
#endif

namespace RSRX_NS {

constexpr int NB = RSRX_MAXBODY;  // 16
constexpr int NJ = RSRX_MAXJNT;   // 12
constexpr int NQ = RSRX_MAXQ;     // 24
constexpr int NV = RSRX_MAXV;     // 20
constexpr int NU = RSRX_MAXU;     // 8
constexpr int NG = RSRX_MAXGEOM;  // 32
constexpr int NS = RSRX_MAXSITE;  // 4
constexpr int NP = RSRX_MAXPAIR;  // 64
#ifndef RSRX_MAXC
#define RSRX_MAXC 24
#endif
constexpr int MAXC = RSRX_MAXC;          // active-contact cap per env (overflow -> status bit)
constexpr int MAXSR = 20;         // sparse rows: equality + dof friction + joint limits
constexpr int MAXROW = MAXSR + 6 * MAXC;
constexpr int LD = NV + 1;        // padded leading dimension of the dense nv x nv matrices
constexpr int NTRI = NV * (NV + 1) / 2;
__host__ __device__ constexpr int tri(int p) { return (p * (p + 1)) >> 1; }  // start of row p in a packed lower triangle
constexpr int MAXMENT = 96;       // (i, ancestor j) entries of the mass matrix
constexpr int MAXTREE_ = 4;
constexpr int OBS_STRIDE = 24;
constexpr int METRICS_STRIDE = 8;

constexpr float MJ_MINVAL = 1e-15f;
constexpr float MJ_MINIMP = 0.0001f;
constexpr float MJ_MAXIMP = 0.9999f;

// Device model: float32 narrowing of rsrx_model_blob plus host-precomputed
// static products (static body poses, static geom poses, mixed pair parameters,
// ancestor masks, mass-matrix entry list).  Lives in global memory; every warp
// reads the same words, so it stays L1-resident.
struct DModel {
  int nbody, njnt, nq, nv, nu, ngeom, nsite, npair, neq, nlevel, nment, ntri;
  int iterations, ls_iterations;
  float timestep, gravity[3], tolerance, ls_tolerance, impratio, meaninertia;
  // bodies
  int body_parentid[NB], body_rootid[NB], body_jntadr[NB], body_jntnum[NB], body_dofadr[NB], body_dofnum[NB],
      body_depth[NB], body_static[NB], body_subtree_end[NB];
  uint32_t body_dofmask[NB];  // dofs on the path from the body to the world
  int body_treeid[NB], ntree, tree_dofadr[MAXTREE_], tree_dofnum[MAXTREE_];
  int dof_tree_lo[NV], dof_tree_hi[NV];  // dof range of the kinematic tree a dof belongs to (M is block diagonal over trees)
  float body_pos[NB][3], body_quat[NB][4], body_ipos[NB][3], body_iquat[NB][4], body_mass[NB], body_inertia[NB][3],
      body_invweight0[NB];
  float static_xpos[NB][3], static_xquat[NB][4];
  // joints
  int jnt_type[NJ], jnt_qposadr[NJ], jnt_dofadr[NJ], jnt_bodyid[NJ], jnt_limited[NJ];
  float jnt_pos[NJ][3], jnt_axis[NJ][3], jnt_range[NJ][2], jnt_solref[NJ][2], jnt_solimp[NJ][5], jnt_margin[NJ];
  float qpos0[NQ];
  // dofs
  int dof_bodyid[NV], dof_jntid[NV], dof_parentid[NV], dof_actfrclimited[NV], dof_hasfriction[NV];
  float dof_damping[NV], dof_frictionloss[NV], dof_armature[NV], dof_invweight0[NV], dof_solref[NV][2],
      dof_solimp[NV][5], dof_actfrcrange[NV][2];
  unsigned char ment_i[MAXMENT], ment_j[MAXMENT];
  unsigned char tri_i[NTRI], tri_j[NTRI];
  // block permutation for the Cholesky factorisations + structurally non-zero entries of H
  int pos_of_dof[NV], dof_of_pos[NV], blk_start[NV], blk_end[NV], tblk_start[NV], tblk_end[NV], nhent;
  int blk_max, tblk_max;  // size of the largest block / tree block
  unsigned char hent_i[NTRI], hent_j[NTRI];
  unsigned short tri_src[NTRI], tri_dst[NTRI];  // M[i][j] offset -> permuted H offset, per lower-triangle entry
  // geoms
  int geom_type[NG], geom_bodyid[NG], geom_static[NG];
  float geom_pos[NG][3], geom_quat[NG][4], geom_size[NG][3], geom_friction[NG][3];
  float geom_static_xpos[NG][3], geom_static_xmat[NG][9];
  float geom_rbound[NG];  // radius of the bounding sphere (|size| of a box)
  // sites
  int site_bodyid[NS];
  float site_pos[NS][3];
  // collision pairs, static mixing precomputed (collision_driver.collision)
  int pair_g1[NP], pair_g2[NP];
  float pair_solref[NP][2], pair_solimp[NP][5], pair_margin[NP], pair_tran[NP];
  // actuators
  int act_qadr[NU], act_dof[NU], act_ctrllimited[NU], act_forcelimited[NU];
  float act_gear[NU], act_gain[NU], act_bias[NU][3], act_ctrlrange[NU][2], act_forcerange[NU][2];
  // joint equality
  int eq_q1[RSRX_MAXEQ], eq_q2[RSRX_MAXEQ], eq_d1[RSRX_MAXEQ], eq_d2[RSRX_MAXEQ];
  float eq_data[RSRX_MAXEQ][5], eq_solref[RSRX_MAXEQ][2], eq_solimp[RSRX_MAXEQ][5], eq_invweight[RSRX_MAXEQ];
  // env
  int env_kind, episode_length, action_repeat, n_frames, cube_body, target_body, site_endpoint, site_tail,
      site_target_tail, geom_base, geom_vertical, geom_target_base, geom_target_vertical, joint_qadr[6];
  float action_scale[NU], push_reward_weight, siet_to_box_reward_weight, healthy_reward, endpoint_min_z_pos;
  // layout
  rsrx_layout lay;
  // shared-memory arena: floats per env (ar::FIXED + pool_floats) and the size of the Jacobian-row pool at its end
  int arena_stride, pool_floats;
  // an env-step that sees more than contact_cap (<= MAXC) active contacts is not committed but handed to the
  // large-capacity kernel (RSRX_STATUS_CONTACT_REDO); MAXC unless a test lowers it (RSRX_CONTACT_CAP)
  int contact_cap;
};

// ---- per-warp shared-memory arena (float words) --------------------------------
// 16.4 KB per env so that 14 envs (2 CTAs x 7 warps) are resident per SM.  Arrays
// that only live before the solver (region P) share storage with arrays that only
// live inside it (region S); mjx's fwd_velocity stages therefore run before
// collision / make_constraint (same results).
constexpr int MAXTREE = MAXTREE_;   // kinematic trees (top-level bodies)
constexpr int NCOL = 14;     // widest contact Jacobian: dofs of the two trees a collision pair joins
// envs (= warps) per CTA at most: the arena is sized so that this many fill one SM's shared memory
#ifndef RSRX_WPB
#define RSRX_WPB 19
#endif
namespace ar {
constexpr int PTRS = 0;  // 4 device pointers (per-env geom_friction, body_mass, dof_frictionloss, Jacobian-row spill) + 2 int flags
constexpr int FLAGS = 8; // int index into PTRS: [8] = "a contact couples two kinematic trees"
constexpr int QPOS = PTRS + 12;
constexpr int QVEL = QPOS + NQ;
constexpr int CTRL = QVEL + NV;
constexpr int WARM = CTRL + NU;
constexpr int DAMP = WARM + NV;
constexpr int XPOS = DAMP + NV;
constexpr int XQUAT = XPOS + NB * 3;
constexpr int SXPOS = XQUAT + NB * 4;
constexpr int SCOM = SXPOS + NS * 3;       // subtree com of each kinematic tree root
constexpr int CDOF = SCOM + MAXTREE * 3;
constexpr int MM = CDOF + NV * 6;          // lower triangle, packed: M[i][j] at i (i + 1) / 2 + j
constexpr int HH = MM + NTRI;              // lower triangle, packed by rows (row p at tri(p)), block-permuted order; also scratch
// nv-vectors
constexpr int V_SMOOTH = HH + NTRI;
constexpr int V_QACCS = V_SMOOTH + NV;
constexpr int V_QACC = V_QACCS + NV;
constexpr int V_QFRCC = V_QACC + NV;
constexpr int V_MA = V_QFRCC + NV;
constexpr int V_GRAD = V_MA + NV;
constexpr int V_MGRAD = V_GRAD + NV;
constexpr int V_SEARCH = V_MGRAD;     // search = -Mgrad, negated in place
constexpr int V_MV = V_MGRAD + NV;
constexpr int V_TMP = V_MV + NV;
constexpr int V_RDIAG = V_TMP + NV;  // 1 / L_kk of the factor in HH
constexpr int V_ACT = V_TMP;          // qfrc_actuator (debug dump only) shares V_TMP (integration scratch)
// contacts: the part of the record that lives through the solver (cf::)
constexpr int CSTRIDE = 10;
constexpr int CON = V_RDIAG + NV;
// efc rows
constexpr int E_AREF = CON + MAXC * CSTRIDE;
constexpr int E_DS = E_AREF + MAXROW;     // D of the sparse rows (contact rows: D in the contact record)
// sparse-row meta
constexpr int SR_DOFA = E_DS + MAXSR;     // int
constexpr int SR_DOFB = SR_DOFA + MAXSR;  // int (-1 = none)
constexpr int SR_CA = SR_DOFB + MAXSR;
constexpr int SR_CB = SR_CA + MAXSR;
constexpr int SR_FLOSS = SR_CB + MAXSR;
constexpr int SR_TYPE = SR_FLOSS + MAXSR;  // int: 0 equality, 1 friction, 2 limit
constexpr int SR_RF = SR_TYPE + MAXSR;     // R * frictionloss
constexpr int OBSBUF = SR_RF + MAXSR;
constexpr int V_BIAS = OBSBUF;             // qfrc_bias (debug dump only) shares the obs staging buffer
// ---- union: region P (alive until velocity_and_forces is done) ...
constexpr int U = OBSBUF + OBS_STRIDE;
constexpr int XANCHOR = U;
constexpr int XAXIS = XANCHOR + NJ * 3;
constexpr int CINERT = XAXIS + NJ * 3;
constexpr int CRB = CINERT + NB * 10;     // crb while building M, then cacc
constexpr int CDOFDOT = CRB + NB * 10;    // crb_cdof while building M, then cdof_dot
constexpr int CVEL = CDOFDOT + NV * 6;
constexpr int P_END = CVEL + NB * 6;
// ---- ... and region S (alive from make_constraint's contact pass to the end of the solver)
constexpr int E_JAREF = U;
constexpr int E_JV = E_JAREF + MAXROW;
constexpr int E_ACT = E_JV + MAXROW;       // one bit per row
constexpr int UB = E_ACT + (MAXROW + 31) / 32;  // [c][4] base-row scratch ...
constexpr int CW = UB;                     // ... / [c][8] per-contact Hessian weights (never live together)
constexpr int S_END = CW + MAXC * 8;
// the transient part of the contact record (ct::: position + frame) is only read while the Jacobian rows are built, i.e.
// before the solver first writes E_JAREF / E_JV
constexpr int CTMP = E_JAREF;              // [c][CTSTRIDE]
constexpr int CTSTRIDE = 12;
static_assert(MAXC * CTSTRIDE <= 2 * MAXROW, "transient contact data must fit E_JAREF + E_JV");
constexpr int U_END = (P_END > S_END ? P_END : S_END);
// ---- Jacobian base rows: a POOL at the end of the arena, DModel::pool_floats long (chosen at model creation so that
// the arena fills the SM's shared memory at WPB envs per CTA).  Contact c owns 4 rows of width na + nb at cf::BOFF; what
// does not fit goes to the env's global-memory spill row (cf::BOFF < 0).  The arena stride is FIXED + pool_floats.
constexpr int BROW = U_END;
constexpr int FIXED = BROW;
// ---- region C (collision only: the pool is empty then): world poses of all geoms + surviving pairs + clip scratch
constexpr int GPOSE = BROW;                // [g][12]: pos(3), mat(9)
constexpr int PLIST = GPOSE + NG * 12;     // int [MAXPAIR]
constexpr int CSCRATCH = PLIST + RSRX_MAXPAIR;  // 8 clipped polygon vertices per half warp
constexpr int C_END = CSCRATCH + 2 * 24;
constexpr int MIN_POOL = C_END - BROW;     // the pool must at least hold region C
constexpr int SPILL_STRIDE = MAXC * 4 * NCOL;  // floats per env in the global spill buffer (contact c at c * 4 * NCOL)
}  // namespace ar

// contact record fields
namespace cf {
constexpr int DIST = 0, MU = 1, KIMPD = 4, B = 5, D = 6, BODIES = 7, COLS = 8, BOFF = 9;  // MU: mu, mu, torsion
// COLS packs the dof ranges of the two kinematic trees the contact joins: a0 | na << 8 | b0 << 16 | nb << 24;
// Jacobian column c is dof a0 + c (c < na) or b0 + c - na.  BOFF: int offset of the contact's 4 x (na + nb) base rows
// in the pool, or -(spill offset + 1).
}
// transient contact fields (ar::CTMP)
namespace ct {
constexpr int POS = 0, FRAME = 3;
}

// ---- debug dump layout (floats per env) -----------------------------------------
namespace dbg {
constexpr int M = 0;                    // nv*nv (row-major, nv = model nv)
constexpr int BIAS = M + NV * NV;
constexpr int QACC_SMOOTH = BIAS + NV;
constexpr int QACC = QACC_SMOOTH + NV;
constexpr int QFRC_C = QACC + NV;
constexpr int NCON = QFRC_C + NV;
constexpr int NEFC = NCON + 1;
constexpr int NITER = NEFC + 1;
constexpr int CDIST = NITER + 1;        // MAXC
constexpr int CPOS = CDIST + MAXC;      // MAXC*3
constexpr int CGEOM = CPOS + MAXC * 3;  // MAXC (g1*64+g2)
constexpr int QFRC_ACT = CGEOM + MAXC;
constexpr int LS_TOTAL = QFRC_ACT + NV;
constexpr int STRIDE = LS_TOTAL + 1;
}  // namespace dbg

// ---- small math (mirrors mjx/_src/math.py) ---------------------------------------
__device__ __forceinline__ float dot3(const float* a, const float* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
__device__ __forceinline__ void cross3(float* r, const float* a, const float* b) {
  float x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void normalize3(float* a) {
  float s = sqrtf(dot3(a, a));
  if (s == 0.f) return;
  a[0] /= s; a[1] /= s; a[2] /= s;
}
__device__ __forceinline__ void normalize4(float* a) {
  float s = sqrtf(a[0] * a[0] + a[1] * a[1] + a[2] * a[2] + a[3] * a[3]);
  if (s == 0.f) return;
  a[0] /= s; a[1] /= s; a[2] /= s; a[3] /= s;
}
__device__ __forceinline__ void quat_mul(float* r, const float* u, const float* v) {
  float a = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  float b = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  float c = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  float d = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = a; r[1] = b; r[2] = c; r[3] = d;
}
__device__ __forceinline__ void rotate(float* r, const float* v, const float* q) {
  float s = q[0];
  const float* u = q + 1;
  float ud = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(c, u, v);
  float x = 2.f * ud * u[0] + (s * s - uu) * v[0] + 2.f * s * c[0];
  float y = 2.f * ud * u[1] + (s * s - uu) * v[1] + 2.f * s * c[1];
  float z = 2.f * ud * u[2] + (s * s - uu) * v[2] + 2.f * s * c[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void quat_to_mat(float* m, const float* q) {
  float q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  float q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  float q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2.f * (q12 - q03); m[2] = 2.f * (q13 + q02);
  m[3] = 2.f * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2.f * (q23 - q01);
  m[6] = 2.f * (q13 - q02); m[7] = 2.f * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
__device__ __forceinline__ void mat_vec(float* r, const float* m, const float* v) {
  float x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  float y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  float z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void matT_vec(float* r, const float* m, const float* v) {
  float x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  float y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  float z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
__device__ __forceinline__ void inert_mul(float* r, const float* i, const float* v) {
  float c1[3], c2[3];
  const float* pos = i + 6;
  cross3(c1, pos, v + 3);
  cross3(c2, pos, v);
  float r0 = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  float r1 = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  float r2 = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  float r3 = i[9] * v[3] - c2[0], r4 = i[9] * v[4] - c2[1], r5 = i[9] * v[5] - c2[2];
  r[0] = r0; r[1] = r1; r[2] = r2; r[3] = r3; r[4] = r4; r[5] = r5;
}
__device__ __forceinline__ void motion_cross(float* r, const float* u, const float* v) {
  float a[3], b[3], c[3];
  cross3(a, u, v);
  cross3(b, u, v + 3);
  cross3(c, u + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
__device__ __forceinline__ void motion_cross_force(float* r, const float* v, const float* f) {
  float a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
__device__ __forceinline__ void make_frame(float* f, const float* n) {
  float a[3] = {n[0], n[1], n[2]};
  normalize3(a);
  float b[3] = {0.f, 0.f, 0.f};
  if (a[1] > -0.5f && a[1] < 0.5f) b[1] = 1.f; else b[2] = 1.f;
  float d = dot3(a, b);
  b[0] -= a[0] * d; b[1] -= a[1] * d; b[2] -= a[2] * d;
  normalize3(b);
  float c[3];
  cross3(c, a, b);
  f[0] = a[0]; f[1] = a[1]; f[2] = a[2];
  f[3] = b[0]; f[4] = b[1]; f[5] = b[2];
  f[6] = c[0]; f[7] = c[1]; f[8] = c[2];
}
__device__ __noinline__ float pw(float x, float p) {  // (noinline: four inlined powf expansions were 6 KB of the kernel)
  if (p == 1.f) return x;
  if (p == 2.f) return x * x;
  return powf(x, p);
}
__device__ __forceinline__ float clipf(float x, float lo, float hi) { return x < lo ? lo : (x > hi ? hi : x); }
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

}  // namespace RSRX_NS
