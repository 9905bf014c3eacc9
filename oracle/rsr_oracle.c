/* rsr_oracle.c — CPU ORACLE.  TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Only tests/ (and the parity / golden-vector scripts under tools/ that serve them:
 * make_golden, dev_gpu_check, solver_stats), __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path (rsr_mjx_b200/ + librsrx.so)
 * never calls into it.
 *
 * PARITY UNPINNED: the arithmetic of the reference's hot path lives in
 * un-vendored pip dependencies (mujoco-mjx==3.2.4, mujoco==3.2.4, brax==0.12.1 on
 * jax==0.4.29, reference README.md:40-57) that are not installed here and
 * cannot be fetched; the reference ships no tests / golden vectors (SURVEY §4,
 * §8c).  This file restates the published MJX 3.2.4 algorithm stage by stage
 * from the upstream sources' structure ([upstream] tags name the file/function
 * restated) and follows the reference's own env code line by line
 * ([ref] tags give file:line under /root/reference).
 *
 * Scalar C, one env at a time, precision selected by -DORACLE_F32 (float) or
 * default double.  oracle_rollout() adds an OpenMP loop over envs for the CPU
 * baseline.  `dense` != 0 reproduces MJX's work pattern (every geom pair emits
 * 4 contact slots, every slot 6 efc rows, inactive rows zeroed); dense == 0
 * drops rows that MJX zeroes.  Both give the same numbers.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "rsr_oracle.h"

#ifdef ORACLE_F32
typedef float real;
#define R_SQRT sqrtf
#define R_SIN sinf
#define R_COS cosf
#define R_ATAN2 atan2f
#define R_TANH tanhf
#define R_ACOS acosf
#define R_POW powf
#define R_FABS fabsf
#else
typedef double real;
#define R_SQRT sqrt
#define R_SIN sin
#define R_COS cos
#define R_ATAN2 atan2
#define R_TANH tanh
#define R_ACOS acos
#define R_POW pow
#define R_FABS fabs
#endif

#define NB RSRX_MAXBODY
#define NJ RSRX_MAXJNT
#define NQ RSRX_MAXQ
#define NV RSRX_MAXV
#define NU RSRX_MAXU
#define NG RSRX_MAXGEOM
#define NS RSRX_MAXSITE
#define MJ_MINVAL ((real)1e-15)
#define MJ_MINIMP ((real)0.0001)
#define MJ_MAXIMP ((real)0.9999)

typedef struct {
  real dist, pos[3], frame[9], friction[5], solref[2], solimp[5];
  int g1, g2;
} contact_t;

typedef struct work {
  const rsrx_model_blob* m;
  int dense;
  /* per-env model arrays (domain randomisation: [ref] domain_randomize.py:26-91) */
  real geom_friction[NG][3], body_mass[NB], dof_damping[NV], dof_frictionloss[NV];
  /* state */
  real qpos[NQ], qvel[NV], ctrl[NU], warm[NV], time;
  /* position-dependent */
  real xpos[NB][3], xquat[NB][4], xmat[NB][9], xipos[NB][3], ximat[NB][9];
  real xanchor[NJ][3], xaxis[NJ][3];
  real gxpos[NG][3], gxmat[NG][9], sxpos[NS][3];
  real subtree_com[NB][3], cinert[NB][10], crb[NB][10], cdof[NV][6];
  real M[NV][NV], L[NV][NV];
  /* velocity-dependent */
  real cvel[NB][6], cdof_dot[NV][6];
  real qfrc_bias[NV], qfrc_passive[NV], qfrc_actuator[NV], qfrc_smooth[NV];
  real qacc_smooth[NV], qacc[NV], qfrc_constraint[NV];
  /* contacts + constraints */
  int ncon, ncon_active;
  contact_t con[ORC_MAXCON];
  int nefc, ne, nf, nefc_active;
  real J[ORC_MAXEFC][NV], D[ORC_MAXEFC], aref[ORC_MAXEFC], floss[ORC_MAXEFC],
      force[ORC_MAXEFC];
  /* solver scratch */
  real Jaref[ORC_MAXEFC], Jv[ORC_MAXEFC], quad[ORC_MAXEFC][3];
  unsigned char active[ORC_MAXEFC];
  int solver_niter, ls_total;
} work;

/* ------------------------------------------------------------------ small math
 * [upstream] mujoco/mjx/_src/math.py */
static inline real dot3(const real* a, const real* b) { return a[0] * b[0] + a[1] * b[1] + a[2] * b[2]; }
static inline void cross3(real* r, const real* a, const real* b) {
  real x = a[1] * b[2] - a[2] * b[1], y = a[2] * b[0] - a[0] * b[2], z = a[0] * b[1] - a[1] * b[0];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline real norm3(const real* a) { return R_SQRT(dot3(a, a)); }
static inline void normalize_n(real* a, int n) {
  real s = 0;
  for (int i = 0; i < n; i++) s += a[i] * a[i];
  s = R_SQRT(s);
  if (s == 0) return; /* math.normalize returns 0 for a zero vector */
  for (int i = 0; i < n; i++) a[i] /= s;
}
static inline void quat_mul(real* r, const real* u, const real* v) {
  real a = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  real b = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  real c = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  real d = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = a; r[1] = b; r[2] = c; r[3] = d;
}
/* math.rotate(vec, quat) */
static inline void rotate(real* r, const real* v, const real* q) {
  real s = q[0];
  const real* u = q + 1;
  real ud = dot3(u, v), uu = dot3(u, u), c[3];
  cross3(c, u, v);
  for (int i = 0; i < 3; i++) r[i] = 2 * ud * u[i] + (s * s - uu) * v[i] + 2 * s * c[i];
}
static inline void quat_to_mat(real* m, const real* q) {
  real q00 = q[0] * q[0], q01 = q[0] * q[1], q02 = q[0] * q[2], q03 = q[0] * q[3];
  real q11 = q[1] * q[1], q12 = q[1] * q[2], q13 = q[1] * q[3];
  real q22 = q[2] * q[2], q23 = q[2] * q[3], q33 = q[3] * q[3];
  m[0] = q00 + q11 - q22 - q33; m[1] = 2 * (q12 - q03); m[2] = 2 * (q13 + q02);
  m[3] = 2 * (q12 + q03); m[4] = q00 - q11 + q22 - q33; m[5] = 2 * (q23 - q01);
  m[6] = 2 * (q13 - q02); m[7] = 2 * (q23 + q01); m[8] = q00 - q11 - q22 + q33;
}
static inline void mat_vec(real* r, const real* m, const real* v) { /* r = M v */
  real x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2];
  real y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2];
  real z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
static inline void matT_vec(real* r, const real* m, const real* v) { /* r = M^T v */
  real x = m[0] * v[0] + m[3] * v[1] + m[6] * v[2];
  real y = m[1] * v[0] + m[4] * v[1] + m[7] * v[2];
  real z = m[2] * v[0] + m[5] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
/* math.inert_mul: cinert = [Ixx Iyy Izz Ixy Ixz Iyz, m*off(3), m] */
static inline void inert_mul(real* r, const real* i, const real* v) {
  real c1[3], c2[3];
  const real* pos = i + 6;
  cross3(c1, pos, v + 3);
  cross3(c2, pos, v);
  r[0] = i[0] * v[0] + i[3] * v[1] + i[4] * v[2] + c1[0];
  r[1] = i[3] * v[0] + i[1] * v[1] + i[5] * v[2] + c1[1];
  r[2] = i[4] * v[0] + i[5] * v[1] + i[2] * v[2] + c1[2];
  r[3] = i[9] * v[3] - c2[0];
  r[4] = i[9] * v[4] - c2[1];
  r[5] = i[9] * v[5] - c2[2];
}
static inline void motion_cross(real* r, const real* u, const real* v) {
  real a[3], b[3], c[3];
  cross3(a, u, v);
  cross3(b, u, v + 3);
  cross3(c, u + 3, v);
  r[0] = a[0]; r[1] = a[1]; r[2] = a[2];
  r[3] = b[0] + c[0]; r[4] = b[1] + c[1]; r[5] = b[2] + c[2];
}
static inline void motion_cross_force(real* r, const real* v, const real* f) {
  real a[3], b[3], c[3];
  cross3(a, v, f);
  cross3(b, v + 3, f + 3);
  cross3(c, v, f + 3);
  r[0] = a[0] + b[0]; r[1] = a[1] + b[1]; r[2] = a[2] + b[2];
  r[3] = c[0]; r[4] = c[1]; r[5] = c[2];
}
/* math.make_frame: normal -> right-handed frame rows [a, b, a x b] */
static void make_frame(real* f, const real* n) {
  real a[3] = {n[0], n[1], n[2]};
  normalize_n(a, 3);
  real b[3] = {0, 0, 0};
  if (a[1] > (real)-0.5 && a[1] < (real)0.5) b[1] = 1; else b[2] = 1;
  real d = dot3(a, b);
  for (int i = 0; i < 3; i++) b[i] -= a[i] * d;
  normalize_n(b, 3);
  real c[3];
  cross3(c, a, b);
  for (int i = 0; i < 3; i++) { f[i] = a[i]; f[3 + i] = b[i]; f[6 + i] = c[i]; }
}
static inline real pw(real x, real p) { /* x**p, exact for the powers MuJoCo defaults to */
  if (p == 1) return x;
  if (p == 2) return x * x;
  return R_POW(x, p);
}
static inline real clipr(real x, real lo, real hi) { return x < lo ? lo : (x > hi ? hi : x); }

/* dense Cholesky A = L L^T (lower), n x n with leading dimension NV */
static void cholesky(real L[NV][NV], real A[NV][NV], int n) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j <= i; j++) {
      real s = A[i][j];
      for (int k = 0; k < j; k++) s -= L[i][k] * L[j][k];
      if (i == j) L[i][i] = R_SQRT(s > MJ_MINVAL ? s : MJ_MINVAL);
      else L[i][j] = s / L[j][j];
    }
  for (int i = 0; i < n; i++)
    for (int j = i + 1; j < n; j++) L[i][j] = 0;
}
static void chol_solve(real L[NV][NV], int n, real* x) {
  for (int i = 0; i < n; i++) {
    real s = x[i];
    for (int k = 0; k < i; k++) s -= L[i][k] * x[k];
    x[i] = s / L[i][i];
  }
  for (int i = n - 1; i >= 0; i--) {
    real s = x[i];
    for (int k = i + 1; k < n; k++) s -= L[k][i] * x[k];
    x[i] = s / L[i][i];
  }
}

/* ------------------------------------------------------------------ kinematics
 * [upstream] mjx/_src/smooth.py::kinematics */
static void kinematics(work* w) {
  const rsrx_model_blob* m = w->m;
  for (int i = 0; i < 3; i++) w->xpos[0][i] = 0;
  w->xquat[0][0] = 1; w->xquat[0][1] = w->xquat[0][2] = w->xquat[0][3] = 0;
  quat_to_mat(w->xmat[0], w->xquat[0]);
  for (int b = 1; b < m->nbody; b++) {
    int p = m->body_parentid[b];
    real bp[3] = {(real)m->body_pos[b][0], (real)m->body_pos[b][1], (real)m->body_pos[b][2]};
    real bq[4] = {(real)m->body_quat[b][0], (real)m->body_quat[b][1], (real)m->body_quat[b][2], (real)m->body_quat[b][3]};
    real pos[3], quat[4], t[3];
    rotate(t, bp, w->xquat[p]);
    for (int i = 0; i < 3; i++) pos[i] = w->xpos[p][i] + t[i];
    quat_mul(quat, w->xquat[p], bq);
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, qa = m->jnt_qposadr[j];
      real jp[3] = {(real)m->jnt_pos[j][0], (real)m->jnt_pos[j][1], (real)m->jnt_pos[j][2]};
      real ja[3] = {(real)m->jnt_axis[j][0], (real)m->jnt_axis[j][1], (real)m->jnt_axis[j][2]};
      if (m->jnt_type[j] == RSRX_JNT_FREE) {
        for (int i = 0; i < 3; i++) { w->xanchor[j][i] = w->qpos[qa + i]; w->xaxis[j][i] = (i == 2); }
        for (int i = 0; i < 3; i++) pos[i] = w->qpos[qa + i];
        for (int i = 0; i < 4; i++) quat[i] = w->qpos[qa + 3 + i];
        normalize_n(quat, 4);
        for (int i = 0; i < 4; i++) w->qpos[qa + 3 + i] = quat[i]; /* "also normalize qpos" */
      } else {
        real anchor[3], axis[3];
        rotate(anchor, jp, quat);
        for (int i = 0; i < 3; i++) anchor[i] += pos[i];
        rotate(axis, ja, quat);
        for (int i = 0; i < 3; i++) { w->xanchor[j][i] = anchor[i]; w->xaxis[j][i] = axis[i]; }
        real dq = w->qpos[qa] - (real)m->qpos0[qa];
        if (m->jnt_type[j] == RSRX_JNT_HINGE) {
          real s = R_SIN(dq * (real)0.5), c = R_COS(dq * (real)0.5);
          real ql[4] = {c, ja[0] * s, ja[1] * s, ja[2] * s}, q2[4];
          quat_mul(q2, quat, ql);
          for (int i = 0; i < 4; i++) quat[i] = q2[i];
          rotate(t, jp, quat); /* correct for off-centre rotation */
          for (int i = 0; i < 3; i++) pos[i] = anchor[i] - t[i];
        } else {
          for (int i = 0; i < 3; i++) pos[i] += axis[i] * dq;
        }
      }
    }
    for (int i = 0; i < 3; i++) w->xpos[b][i] = pos[i];
    for (int i = 0; i < 4; i++) w->xquat[b][i] = quat[i];
    quat_to_mat(w->xmat[b], quat);
  }
  for (int b = 0; b < m->nbody; b++) {
    real ip[3] = {(real)m->body_ipos[b][0], (real)m->body_ipos[b][1], (real)m->body_ipos[b][2]};
    real iq[4] = {(real)m->body_iquat[b][0], (real)m->body_iquat[b][1], (real)m->body_iquat[b][2], (real)m->body_iquat[b][3]};
    real t[3], q[4];
    rotate(t, ip, w->xquat[b]);
    for (int i = 0; i < 3; i++) w->xipos[b][i] = w->xpos[b][i] + t[i];
    quat_mul(q, w->xquat[b], iq);
    quat_to_mat(w->ximat[b], q);
  }
  for (int g = 0; g < m->ngeom; g++) {
    int b = m->geom_bodyid[g];
    real gp[3] = {(real)m->geom_pos[g][0], (real)m->geom_pos[g][1], (real)m->geom_pos[g][2]};
    real gq[4] = {(real)m->geom_quat[g][0], (real)m->geom_quat[g][1], (real)m->geom_quat[g][2], (real)m->geom_quat[g][3]};
    real t[3], q[4];
    rotate(t, gp, w->xquat[b]);
    for (int i = 0; i < 3; i++) w->gxpos[g][i] = w->xpos[b][i] + t[i];
    quat_mul(q, w->xquat[b], gq);
    quat_to_mat(w->gxmat[g], q);
  }
  for (int s = 0; s < m->nsite; s++) {
    int b = m->site_bodyid[s];
    real sp[3] = {(real)m->site_pos[s][0], (real)m->site_pos[s][1], (real)m->site_pos[s][2]};
    real t[3];
    rotate(t, sp, w->xquat[b]);
    for (int i = 0; i < 3; i++) w->sxpos[s][i] = w->xpos[b][i] + t[i];
  }
}

/* [upstream] smooth.py::com_pos */
static void com_pos(work* w) {
  const rsrx_model_blob* m = w->m;
  real pos[NB][3], mass[NB];
  for (int b = 0; b < m->nbody; b++) {
    mass[b] = w->body_mass[b];
    for (int i = 0; i < 3; i++) pos[b][i] = w->xipos[b][i] * mass[b];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    mass[p] += mass[b];
    for (int i = 0; i < 3; i++) pos[p][i] += pos[b][i];
  }
  for (int b = 0; b < m->nbody; b++)
    for (int i = 0; i < 3; i++)
      w->subtree_com[b][i] = mass[b] < MJ_MINVAL ? w->xipos[b][i] : pos[b][i] / (mass[b] > MJ_MINVAL ? mass[b] : MJ_MINVAL);
  for (int b = 0; b < m->nbody; b++) {
    const real* root = w->subtree_com[m->body_rootid[b]];
    real off[3] = {w->xipos[b][0] - root[0], w->xipos[b][1] - root[1], w->xipos[b][2] - root[2]};
    real ms = w->body_mass[b];
    const real* R = w->ximat[b];
    real I[3] = {(real)m->body_inertia[b][0], (real)m->body_inertia[b][1], (real)m->body_inertia[b][2]};
    real in[3][3];
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) {
        real s = 0;
        for (int k = 0; k < 3; k++) s += R[r * 3 + k] * I[k] * R[c * 3 + k];
        in[r][c] = s;
      }
    /* h = cross(off, -eye(3)); inert += h h^T mass */
    real h[3][3];
    for (int r = 0; r < 3; r++) {
      real e[3] = {0, 0, 0};
      e[r] = -1;
      cross3(h[r], off, e);
    }
    for (int r = 0; r < 3; r++)
      for (int c = 0; c < 3; c++) in[r][c] += dot3(h[r], h[c]) * ms;
    real* ci = w->cinert[b];
    ci[0] = in[0][0]; ci[1] = in[1][1]; ci[2] = in[2][2];
    ci[3] = in[0][1]; ci[4] = in[0][2]; ci[5] = in[1][2];
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
  }
  /* cdof: [angular, linear] motion axes about the tree root's subtree com */
  for (int j = 0; j < m->njnt; j++) {
    int b = m->jnt_bodyid[j], d = m->jnt_dofadr[j];
    const real* root = w->subtree_com[m->body_rootid[b]];
    real off[3] = {root[0] - w->xanchor[j][0], root[1] - w->xanchor[j][1], root[2] - w->xanchor[j][2]};
    if (m->jnt_type[j] == RSRX_JNT_FREE) {
      for (int a = 0; a < 3; a++) {
        for (int i = 0; i < 6; i++) w->cdof[d + a][i] = (i == 3 + a);
        real ax[3] = {w->xmat[b][0 + a], w->xmat[b][3 + a], w->xmat[b][6 + a]}, c[3]; /* xmat.T rows */
        cross3(c, ax, off);
        for (int i = 0; i < 3; i++) { w->cdof[d + 3 + a][i] = ax[i]; w->cdof[d + 3 + a][3 + i] = c[i]; }
      }
    } else if (m->jnt_type[j] == RSRX_JNT_SLIDE) {
      for (int i = 0; i < 3; i++) { w->cdof[d][i] = 0; w->cdof[d][3 + i] = w->xaxis[j][i]; }
    } else {
      real c[3];
      cross3(c, w->xaxis[j], off);
      for (int i = 0; i < 3; i++) { w->cdof[d][i] = w->xaxis[j][i]; w->cdof[d][3 + i] = c[i]; }
    }
  }
}

/* [upstream] smooth.py::crb + support.make_m (dense), factor_m */
static void crb_and_factor(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  memcpy(w->crb, w->cinert, sizeof(w->crb));
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int i = 0; i < 10; i++) w->crb[p][i] += w->crb[b][i];
  }
  for (int i = 0; i < 10; i++) w->crb[0][i] = 0;
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) w->M[i][j] = 0;
  for (int i = 0; i < nv; i++) {
    real f[6];
    inert_mul(f, w->crb[m->dof_bodyid[i]], w->cdof[i]);
    for (int j = i; j >= 0; j = m->dof_parentid[j]) {
      real s = 0;
      for (int k = 0; k < 6; k++) s += w->cdof[j][k] * f[k];
      w->M[i][j] = s;
      w->M[j][i] = s;
    }
    w->M[i][i] += (real)m->dof_armature[i];
  }
  cholesky(w->L, w->M, nv);
}

/* ------------------------------------------------------------------- collision
 * [upstream] mjx/_src/collision_convex.py::_manifold_points */
static void manifold_points(const real (*poly)[3], const unsigned char* mask, int n, const real* nrm, int idx[4]) {
  /* MJX takes four independent argmax'es.  On a rectangle (every box face) the
   * 4th one is an exact three-way tie between the wanted corner and two corners
   * already taken, so XLA's result depends on rounding.  This restatement breaks
   * ties deterministically: a vertex already selected scores like a masked one
   * (-1e6), i.e. selections prefer vertices not yet taken. */
  real dm[32];
  int a = 0, b = 0, c = 0, d = 0;
  real best;
  for (int i = 0; i < n; i++) dm[i] = mask[i] ? (real)0 : (real)-1e6;
  best = dm[0];
  for (int i = 1; i < n; i++) if (dm[i] > best) { best = dm[i]; a = i; }
  dm[a] = (real)-1e6;
  best = -(real)1e30;
  for (int i = 0; i < n; i++) {
    real t[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]};
    real v = dot3(t, t) + dm[i];
    if (v > best) { best = v; b = i; }
  }
  dm[b] = (real)-1e6;
  real ab[3], amb[3] = {poly[a][0] - poly[b][0], poly[a][1] - poly[b][1], poly[a][2] - poly[b][2]};
  cross3(ab, nrm, amb);
  best = -(real)1e30;
  for (int i = 0; i < n; i++) {
    real ap[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]};
    real v = R_FABS(dot3(ap, ab)) + dm[i];
    if (v > best) { best = v; c = i; }
  }
  dm[c] = (real)-1e6;
  real ac[3], bc[3];
  real amc[3] = {poly[a][0] - poly[c][0], poly[a][1] - poly[c][1], poly[a][2] - poly[c][2]};
  real bmc[3] = {poly[b][0] - poly[c][0], poly[b][1] - poly[c][1], poly[b][2] - poly[c][2]};
  cross3(ac, nrm, amc);
  cross3(bc, nrm, bmc);
  best = -(real)1e30;
  for (int i = 0; i < n; i++) { /* concatenate([dist_bp, dist_ap]).argmax() % n */
    real bp[3] = {poly[b][0] - poly[i][0], poly[b][1] - poly[i][1], poly[b][2] - poly[i][2]};
    real v = R_FABS(dot3(bp, bc)) + dm[i];
    if (v > best) { best = v; d = i; }
  }
  for (int i = 0; i < n; i++) {
    real ap[3] = {poly[a][0] - poly[i][0], poly[a][1] - poly[i][1], poly[a][2] - poly[i][2]};
    real v = R_FABS(dot3(ap, ac)) + dm[i];
    if (v > best) { best = v; d = i; }
  }
  idx[0] = a; idx[1] = b; idx[2] = c; idx[3] = d;
}

static void box_vertex(real* v, const real* size, int k) {
  v[0] = (k & 1) ? size[0] : -size[0];
  v[1] = (k & 2) ? size[1] : -size[1];
  v[2] = (k & 4) ? size[2] : -size[2];
}

/* [upstream] collision_convex.py::plane_convex with the 8 box vertices.
 * Returns 4 slots: dist[4], pos[4][3]; normal = plane z axis. */
static void plane_box(const real* ppos, const real* pmat, const real* bpos, const real* bmat, const real* size,
                      real dist[4], real pos[4][3], real nrm[3]) {
  real vert[8][3], support[8], n[3], pp[3], d[3];
  real pn[3] = {pmat[2], pmat[5], pmat[8]};
  for (int i = 0; i < 3; i++) d[i] = ppos[i] - bpos[i];
  matT_vec(pp, bmat, d);
  matT_vec(n, bmat, pn);
  real smax = -(real)1e30;
  for (int k = 0; k < 8; k++) {
    box_vertex(vert[k], size, k);
    real t[3] = {pp[0] - vert[k][0], pp[1] - vert[k][1], pp[2] - vert[k][2]};
    support[k] = dot3(t, n);
    if (support[k] > smax) smax = support[k];
  }
  unsigned char mask[8];
  real thr = smax - (real)1e-3;
  if (thr < 0) thr = 0;
  for (int k = 0; k < 8; k++) mask[k] = support[k] > thr;
  int idx[4];
  manifold_points((const real(*)[3])vert, mask, 8, n, idx);
  for (int c = 0; c < 4; c++) {
    int unique = 1;
    for (int e = 0; e < c; e++) if (idx[e] == idx[c]) unique = 0;
    real wp[3];
    mat_vec(wp, bmat, vert[idx[c]]);
    dist[c] = unique ? -support[idx[c]] : (real)1;
    for (int i = 0; i < 3; i++) pos[c][i] = bpos[i] + wp[i] - (real)0.5 * dist[c] * pn[i];
  }
  for (int i = 0; i < 3; i++) nrm[i] = pn[i];
}

/* Sutherland-Hodgman clip of a convex polygon against the half-space
 * n.x <= h.  Returns the new vertex count. */
static int clip_halfspace(real (*poly)[3], int n, const real* pn, real h, real (*out)[3]) {
  int no = 0;
  for (int i = 0; i < n; i++) {
    const real* a = poly[i];
    const real* b = poly[(i + 1) % n];
    real da = dot3(a, pn) - h, db = dot3(b, pn) - h;
    if (da <= 0) { for (int k = 0; k < 3; k++) out[no][k] = a[k]; no++; }
    if ((da < 0 && db > 0) || (da > 0 && db < 0)) {
      real t = da / (da - db);
      for (int k = 0; k < 3; k++) out[no][k] = a[k] + t * (b[k] - a[k]);
      no++;
    }
  }
  return no;
}

/* box_box: separating-axis test (3+3 face normals, 9 edge crosses) in the frame
 * of box 2, face contact by clipping the incident face against the reference
 * face's side planes + 4-point manifold selection, edge contact by closest
 * points of the two supporting edges.
 * [upstream] collision_convex.py::box_box/_sat/_create_contact_manifold —
 * structure restated from memory; see DESIGN.md "box-box" for the exact rules
 * this oracle fixes (axis order, 1.05 face preference, tie-breaks, dedup).
 * Normal points from box 1 to box 2. */
static void box_box(const real* p1, const real* m1, const real* s1, const real* p2, const real* m2, const real* s2,
                    real dist[4], real pos[4][3], real nrm[3]) {
  real R[9], t[3], d[3];
  for (int i = 0; i < 3; i++) d[i] = p1[i] - p2[i];
  matT_vec(t, m2, d); /* centre of box 1 in frame 2 */
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      real s = 0;
      for (int k = 0; k < 3; k++) s += m2[k * 3 + r] * m1[k * 3 + c];
      R[r * 3 + c] = s; /* R = m2^T m1: columns = box-1 axes in frame 2 */
    }
  real axA[3][3], axB[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
  for (int a = 0; a < 3; a++) for (int i = 0; i < 3; i++) axA[a][i] = R[i * 3 + a];
  real axes[15][3];
  unsigned char degenerate[15];
  for (int a = 0; a < 3; a++) for (int i = 0; i < 3; i++) { axes[a][i] = axA[a][i]; axes[3 + a][i] = axB[a][i]; }
  for (int k = 0; k < 6; k++) degenerate[k] = 0;
  for (int k = 0; k < 9; k++) {
    real c[3];
    cross3(c, axA[k % 3], axB[k / 3]);
    degenerate[6 + k] = dot3(c, c) < (real)1e-6;
    normalize_n(c, 3);
    for (int i = 0; i < 3; i++) axes[6 + k][i] = c[i];
  }
  real overlap[15], sgn[15];
  for (int k = 0; k < 15; k++) {
    const real* ax = axes[k];
    real ca = dot3(t, ax);
    real ra = s1[0] * R_FABS(dot3(axA[0], ax)) + s1[1] * R_FABS(dot3(axA[1], ax)) + s1[2] * R_FABS(dot3(axA[2], ax));
    real rb = s2[0] * R_FABS(ax[0]) + s2[1] * R_FABS(ax[1]) + s2[2] * R_FABS(ax[2]);
    real dist1 = (ca + ra) - (-rb); /* maxA - minB */
    real dist2 = rb - (ca - ra);    /* maxB - minA */
    sgn[k] = dist1 > dist2 ? (real)-1 : (real)1;
    overlap[k] = dist1 < dist2 ? dist1 : dist2;
    if (degenerate[k]) overlap[k] = (real)1e6;
  }
  int bf = 0, be = 6;
  for (int k = 1; k < 6; k++) if (overlap[k] < overlap[bf]) bf = k;
  for (int k = 7; k < 15; k++) if (overlap[k] < overlap[be]) be = k;
  int is_edge = overlap[be] * (real)1.05 < overlap[bf];
  int best = is_edge ? be : bf;
  real n[3] = {axes[best][0] * sgn[best], axes[best][1] * sgn[best], axes[best][2] * sgn[best]}; /* 1 -> 2, frame 2 */
  for (int c = 0; c < 4; c++) { dist[c] = 1; for (int i = 0; i < 3; i++) pos[c][i] = 0; }
  real lp[4][3];
  for (int c = 0; c < 4; c++) for (int i = 0; i < 3; i++) lp[c][i] = 0;

  if (overlap[best] < 0) {
    /* separated along the best axis: MJX marks every slot inactive */
  } else if (!is_edge) {
    /* reference box P (owns the axis), incident box Q */
    int refA = best < 3;
    int r = refA ? best : best - 3;
    const real (*axP)[3] = refA ? axA : axB;
    const real (*axQ)[3] = refA ? axB : axA;
    const real* hP = refA ? s1 : s2;
    const real* hQ = refA ? s2 : s1;
    real cP[3], cQ[3], nref[3];
    for (int i = 0; i < 3; i++) { cP[i] = refA ? t[i] : 0; cQ[i] = refA ? 0 : t[i]; nref[i] = refA ? n[i] : -n[i]; }
    /* incident face: Q's face most anti-parallel to nref */
    int q = 0;
    real bestd = -1;
    for (int k = 0; k < 3; k++) { real v = R_FABS(dot3(axQ[k], nref)); if (v > bestd) { bestd = v; q = k; } }
    real sq = dot3(axQ[q], nref) > 0 ? (real)-1 : (real)1;
    int u = (q + 1) % 3, v = (q + 2) % 3;
    real poly[8][3], tmp[8][3];
    const real su[4] = {1, -1, -1, 1}, sv[4] = {1, 1, -1, -1};
    for (int k = 0; k < 4; k++)
      for (int i = 0; i < 3; i++)
        poly[k][i] = cQ[i] + sq * hQ[q] * axQ[q][i] + su[k] * hQ[u] * axQ[u][i] + sv[k] * hQ[v] * axQ[v][i];
    int np = 4;
    int pu = (r + 1) % 3, pv = (r + 2) % 3;
    const int side_ax[4] = {pu, pu, pv, pv};
    const real side_sg[4] = {1, -1, 1, -1};
    for (int s = 0; s < 4 && np > 0; s++) {
      real pn[3] = {side_sg[s] * axP[side_ax[s]][0], side_sg[s] * axP[side_ax[s]][1], side_sg[s] * axP[side_ax[s]][2]};
      real h = hP[side_ax[s]] + dot3(cP, pn);
      np = clip_halfspace(poly, np, pn, h, tmp);
      memcpy(poly, tmp, sizeof(real) * 3 * np);
    }
    if (np > 0) {
      /* depth below the reference face, projection onto it */
      real depth[8], ref[8][3];
      unsigned char mask[8];
      real sr = dot3(axP[r], nref) > 0 ? (real)1 : (real)-1;
      (void)sr;
      for (int k = 0; k < np; k++) {
        real rel[3] = {poly[k][0] - cP[0], poly[k][1] - cP[1], poly[k][2] - cP[2]};
        depth[k] = hP[r] - dot3(rel, nref);
        mask[k] = depth[k] > 0;
        for (int i = 0; i < 3; i++) ref[k][i] = poly[k][i] + depth[k] * nref[i];
      }
      int idx[4];
      manifold_points((const real(*)[3])ref, mask, np, nref, idx);
      for (int c = 0; c < 4; c++) {
        int unique = 1;
        for (int e = 0; e < c; e++) if (idx[e] == idx[c]) unique = 0;
        int k = idx[c];
        if (unique && mask[k]) {
          dist[c] = -depth[k];
          for (int i = 0; i < 3; i++) lp[c][i] = poly[k][i] + (real)0.5 * depth[k] * nref[i];
        }
      }
    }
  } else {
    /* edge-edge: supporting edges of A (dir axA[i]) and B (dir axB[j]) */
    int ia = (best - 6) % 3, jb = (best - 6) / 3;
    real ea[3], eb[3];
    for (int i = 0; i < 3; i++) { ea[i] = t[i]; eb[i] = 0; }
    for (int k = 0; k < 3; k++) {
      if (k != ia) {
        real s = dot3(axA[k], n) >= 0 ? (real)1 : (real)-1;
        for (int i = 0; i < 3; i++) ea[i] += s * s1[k] * axA[k][i];
      }
      if (k != jb) {
        real s = dot3(axB[k], n) >= 0 ? (real)1 : (real)-1;
        for (int i = 0; i < 3; i++) eb[i] -= s * s2[k] * axB[k][i];
      }
    }
    /* closest points of lines ea + sa*ua, eb + sb*ub, clamped to the edges */
    const real* ua = axA[ia];
    const real* ub = axB[jb];
    real w0[3] = {ea[0] - eb[0], ea[1] - eb[1], ea[2] - eb[2]};
    real bb = dot3(ua, ub), dd = dot3(ua, w0), ee = dot3(ub, w0);
    real den = 1 - bb * bb;
    real sa = den > (real)1e-12 ? (bb * ee - dd) / den : 0;
    real sb = den > (real)1e-12 ? (ee - bb * dd) / den : 0;
    sa = clipr(sa, -s1[ia], s1[ia]);
    sb = clipr(sb, -s2[jb], s2[jb]);
    real pa[3], pb[3], df[3];
    for (int i = 0; i < 3; i++) { pa[i] = ea[i] + sa * ua[i]; pb[i] = eb[i] + sb * ub[i]; df[i] = pb[i] - pa[i]; }
    dist[0] = dot3(df, n);
    for (int i = 0; i < 3; i++) lp[0][i] = (real)0.5 * (pa[i] + pb[i]);
  }
  /* back to the world frame */
  for (int c = 0; c < 4; c++) {
    real wp[3];
    mat_vec(wp, m2, lp[c]);
    for (int i = 0; i < 3; i++) pos[c][i] = p2[i] + wp[i];
  }
  mat_vec(nrm, m2, n);
}

/* [upstream] collision_driver.py::collision (contact parameter mixing) */
static void collision(work* w) {
  const rsrx_model_blob* m = w->m;
  w->ncon = 0;
  w->ncon_active = 0;
  for (int p = 0; p < m->npair; p++) {
    int g1 = m->pair_geom1[p], g2 = m->pair_geom2[p];
    real dist[4], pos[4][3], nrm[3];
    real sz1[3] = {(real)m->geom_size[g1][0], (real)m->geom_size[g1][1], (real)m->geom_size[g1][2]};
    real sz2[3] = {(real)m->geom_size[g2][0], (real)m->geom_size[g2][1], (real)m->geom_size[g2][2]};
    if (m->geom_type[g1] == RSRX_GEOM_PLANE)
      plane_box(w->gxpos[g1], w->gxmat[g1], w->gxpos[g2], w->gxmat[g2], sz2, dist, pos, nrm);
    else
      box_box(w->gxpos[g1], w->gxmat[g1], sz1, w->gxpos[g2], w->gxmat[g2], sz2, dist, pos, nrm);
    /* parameter mixing */
    real mix1 = (real)m->geom_solmix[g1], mix2 = (real)m->geom_solmix[g2];
    real mix = mix1 / (mix1 + mix2);
    if (mix1 < MJ_MINVAL && mix2 < MJ_MINVAL) mix = (real)0.5;
    else if (mix1 < MJ_MINVAL) mix = 0;
    else if (mix2 < MJ_MINVAL) mix = 1;
    real fr[3], solref[2], solimp[5];
    for (int i = 0; i < 3; i++) fr[i] = w->geom_friction[g1][i] > w->geom_friction[g2][i] ? w->geom_friction[g1][i] : w->geom_friction[g2][i];
    real r1[2] = {(real)m->geom_solref[g1][0], (real)m->geom_solref[g1][1]};
    real r2[2] = {(real)m->geom_solref[g2][0], (real)m->geom_solref[g2][1]};
    int standard = r1[0] > 0 && r2[0] > 0;
    for (int i = 0; i < 2; i++) solref[i] = standard ? mix * r1[i] + (1 - mix) * r2[i] : (r1[i] < r2[i] ? r1[i] : r2[i]);
    for (int i = 0; i < 5; i++) solimp[i] = mix * (real)m->geom_solimp[g1][i] + (1 - mix) * (real)m->geom_solimp[g2][i];
    real margin = (real)(m->geom_margin[g1] > m->geom_margin[g2] ? m->geom_margin[g1] : m->geom_margin[g2]);
    real frame[9];
    make_frame(frame, nrm);
    for (int c = 0; c < 4; c++) {
      int act = dist[c] - margin < 0;
      if (act) w->ncon_active++;
      if (!act && !w->dense) continue;
      contact_t* k = &w->con[w->ncon++];
      k->dist = dist[c];
      for (int i = 0; i < 3; i++) k->pos[i] = pos[c][i];
      memcpy(k->frame, frame, sizeof(frame));
      k->friction[0] = fr[0]; k->friction[1] = fr[0]; k->friction[2] = fr[1]; k->friction[3] = fr[2]; k->friction[4] = fr[2];
      k->solref[0] = solref[0]; k->solref[1] = solref[1];
      memcpy(k->solimp, solimp, sizeof(solimp));
      k->g1 = g1; k->g2 = g2;
    }
  }
}

/* ----------------------------------------------------------------- constraints
 * [upstream] mjx/_src/constraint.py::_kbi / _row */
static void kbi(const work* w, const real* solref, const real* solimp, real pos, real* k, real* b, real* imp) {
  real timeconst = solref[0], dampratio = solref[1];
  real dt2 = 2 * (real)w->m->timestep;
  if (timeconst < dt2) timeconst = dt2; /* refsafe */
  real dmin = clipr(solimp[0], MJ_MINIMP, MJ_MAXIMP), dmax = clipr(solimp[1], MJ_MINIMP, MJ_MAXIMP);
  real width = solimp[2] > MJ_MINVAL ? solimp[2] : MJ_MINVAL;
  real mid = clipr(solimp[3], MJ_MINIMP, MJ_MAXIMP);
  real power = solimp[4] > 1 ? solimp[4] : 1;
  *k = 1 / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2 / (dmax * timeconst);
  if (solref[0] <= 0) *k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0) *b = -solref[1] / dmax;
  real x = R_FABS(pos) / width, y;
  if (x > 1) { *imp = dmax; return; }
  if (x < mid) y = ((real)1 / pw(mid, power - 1)) * pw(x, power);
  else y = 1 - ((real)1 / pw(1 - mid, power - 1)) * pw(1 - x, power);
  *imp = clipr(dmin + y * (dmax - dmin), dmin, dmax);
}

static void add_row(work* w, const real* Jrow, real pos, real invweight, const real* solref, const real* solimp,
                    real margin, real frictionloss) {
  int nv = w->m->nv, r = w->nefc++;
  real vel = 0, k, b, imp;
  for (int i = 0; i < nv; i++) { w->J[r][i] = Jrow[i]; vel += Jrow[i] * w->qvel[i]; }
  kbi(w, solref, solimp, pos - margin, &k, &b, &imp);
  real rr = invweight * (1 - imp) / imp;
  if (rr < MJ_MINVAL) rr = MJ_MINVAL;
  w->D[r] = 1 / rr;
  w->aref[r] = -b * vel - k * imp * (pos - margin);
  w->floss[r] = frictionloss;
}

/* [upstream] support.py::jac — (jacp, jacr)[dof] of a world point on a body */
static void jac_point(const work* w, const real* point, int body, real jp[NV][3], real jr[NV][3]) {
  const rsrx_model_blob* m = w->m;
  const real* root = w->subtree_com[m->body_rootid[body]];
  real off[3] = {point[0] - root[0], point[1] - root[1], point[2] - root[2]};
  unsigned char anc[NB];
  memset(anc, 0, sizeof(anc));
  for (int b = body; b > 0; b = m->body_parentid[b]) anc[b] = 1;
  for (int d = 0; d < m->nv; d++) {
    if (anc[m->dof_bodyid[d]]) {
      real c[3];
      cross3(c, w->cdof[d], off);
      for (int i = 0; i < 3; i++) { jp[d][i] = w->cdof[d][3 + i] + c[i]; jr[d][i] = w->cdof[d][i]; }
    } else {
      for (int i = 0; i < 3; i++) jp[d][i] = jr[d][i] = 0;
    }
  }
}

/* [upstream] constraint.py::make_constraint — rows: equality, dof friction,
 * joint limits, pyramidal contacts */
static void make_constraint(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  real Jrow[NV];
  w->nefc = 0;
  /* _instantiate_equality_joint */
  for (int e = 0; e < m->neq; e++) {
    int j1 = m->eq_obj1id[e], j2 = m->eq_obj2id[e];
    int q1 = m->jnt_qposadr[j1], d1 = m->jnt_dofadr[j1];
    real solref[2] = {(real)m->eq_solref[e][0], (real)m->eq_solref[e][1]}, solimp[5];
    for (int i = 0; i < 5; i++) solimp[i] = (real)m->eq_solimp[e][i];
    for (int i = 0; i < nv; i++) Jrow[i] = 0;
    real pos1 = w->qpos[q1] - (real)m->qpos0[q1], invw = (real)m->dof_invweight0[d1], pos;
    if (j2 >= 0) {
      int q2 = m->jnt_qposadr[j2], d2 = m->jnt_dofadr[j2];
      real dif = w->qpos[q2] - (real)m->qpos0[q2];
      real dp[5] = {1, dif, dif * dif, dif * dif * dif, dif * dif * dif * dif};
      real deriv = 0, poly = 0;
      for (int i = 0; i < 5; i++) poly += (real)m->eq_data[e][i] * dp[i];
      for (int i = 1; i < 5; i++) deriv += (real)m->eq_data[e][i] * dp[i - 1] * (real)i;
      Jrow[d2] = -deriv;
      pos = pos1 - poly;
      invw += (real)m->dof_invweight0[d2];
    } else {
      pos = pos1 - (real)m->eq_data[e][0];
    }
    Jrow[d1] = 1;
    add_row(w, Jrow, pos, invw, solref, solimp, 0, 0);
  }
  w->ne = w->nefc;
  /* _instantiate_friction (dof frictionloss) */
  for (int d = 0; d < nv; d++) {
    if (!(m->dof_frictionloss[d] > 0)) continue;
    real solref[2] = {(real)m->dof_solref[d][0], (real)m->dof_solref[d][1]}, solimp[5];
    for (int i = 0; i < 5; i++) solimp[i] = (real)m->dof_solimp[d][i];
    for (int i = 0; i < nv; i++) Jrow[i] = 0;
    Jrow[d] = 1;
    add_row(w, Jrow, 0, (real)m->dof_invweight0[d], solref, solimp, 0, w->dof_frictionloss[d]);
  }
  w->nf = w->nefc - w->ne;
  /* _instantiate_limit_slide_hinge */
  for (int j = 0; j < m->njnt; j++) {
    if (!m->jnt_limited[j] || m->jnt_type[j] == RSRX_JNT_FREE) continue;
    int qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    real q = w->qpos[qa];
    real dmin = q - (real)m->jnt_range[j][0], dmax = (real)m->jnt_range[j][1] - q;
    real margin = (real)m->jnt_margin[j];
    real pos = (dmin < dmax ? dmin : dmax) - margin;
    int act = pos < 0;
    if (!act && !w->dense) continue;
    real solref[2] = {(real)m->jnt_solref[j][0], (real)m->jnt_solref[j][1]}, solimp[5];
    for (int i = 0; i < 5; i++) solimp[i] = (real)m->jnt_solimp[j][i];
    for (int i = 0; i < nv; i++) Jrow[i] = 0;
    Jrow[d] = act ? (dmin < dmax ? (real)1 : (real)-1) : 0;
    add_row(w, Jrow, pos, (real)m->dof_invweight0[d], solref, solimp, 0, 0);
  }
  /* _instantiate_contact: condim 4 pyramid, 6 rows per contact */
  for (int c = 0; c < w->ncon; c++) {
    const contact_t* k = &w->con[c];
    int b1 = m->geom_bodyid[k->g1], b2 = m->geom_bodyid[k->g2];
    int condim1 = m->geom_condim[k->g1], condim2 = m->geom_condim[k->g2];
    int condim = condim1 > condim2 ? condim1 : condim2;
    real jp1[NV][3], jr1[NV][3], jp2[NV][3], jr2[NV][3];
    jac_point(w, k->pos, b1, jp1, jr1);
    jac_point(w, k->pos, b2, jp2, jr2);
    int act = k->dist < 0;
    real tran = (real)m->body_invweight0[b1][0] + (real)m->body_invweight0[b2][0];
    /* Jacobian difference rotated into the contact frame: rows normal, t1, t2, (rot about normal, ...) */
    real diff[6][NV];
    for (int d = 0; d < nv; d++) {
      real dp[3] = {jp2[d][0] - jp1[d][0], jp2[d][1] - jp1[d][1], jp2[d][2] - jp1[d][2]};
      real dr[3] = {jr2[d][0] - jr1[d][0], jr2[d][1] - jr1[d][1], jr2[d][2] - jr1[d][2]};
      for (int a = 0; a < 3; a++) { diff[a][d] = dot3(k->frame + 3 * a, dp); diff[3 + a][d] = dot3(k->frame + 3 * a, dr); }
    }
    real mu0 = k->friction[0];
    real invw = (tran + mu0 * mu0 * tran) * 2 * mu0 * mu0 / (real)m->impratio; /* common to all edges */
    if (condim == 1) {
      for (int d = 0; d < nv; d++) Jrow[d] = act ? diff[0][d] : 0;
      add_row(w, Jrow, k->dist, tran, k->solref, k->solimp, 0, 0);
      continue;
    }
    for (int t = 0; t < condim - 1; t++)
      for (int s = 0; s < 2; s++) {
        real f = s == 0 ? k->friction[t] : -k->friction[t];
        for (int d = 0; d < nv; d++) Jrow[d] = act ? diff[0][d] + diff[1 + t][d] * f : 0;
        add_row(w, Jrow, k->dist, invw, k->solref, k->solimp, 0, 0);
      }
  }
}

/* ------------------------------------------------------------ velocity / forces
 * [upstream] smooth.py::com_vel */
static void com_vel(work* w) {
  const rsrx_model_blob* m = w->m;
  for (int i = 0; i < 6; i++) w->cvel[0][i] = 0;
  for (int b = 1; b < m->nbody; b++) {
    real cvel[6];
    memcpy(cvel, w->cvel[m->body_parentid[b]], sizeof(cvel));
    for (int k = 0; k < m->body_jntnum[b]; k++) {
      int j = m->body_jntadr[b] + k, d = m->jnt_dofadr[j];
      if (m->jnt_type[j] == RSRX_JNT_FREE) {
        for (int a = 0; a < 3; a++)
          for (int i = 0; i < 6; i++) cvel[i] += w->cdof[d + a][i] * w->qvel[d + a];
        for (int a = 0; a < 3; a++) {
          for (int i = 0; i < 6; i++) w->cdof_dot[d + a][i] = 0;
          motion_cross(w->cdof_dot[d + 3 + a], cvel, w->cdof[d + 3 + a]);
        }
        for (int a = 3; a < 6; a++)
          for (int i = 0; i < 6; i++) cvel[i] += w->cdof[d + a][i] * w->qvel[d + a];
      } else {
        motion_cross(w->cdof_dot[d], cvel, w->cdof[d]);
        for (int i = 0; i < 6; i++) cvel[i] += w->cdof[d][i] * w->qvel[d];
      }
    }
    memcpy(w->cvel[b], cvel, sizeof(cvel));
  }
}

/* [upstream] passive.py::passive (damping only), smooth.py::rne */
static void passive_and_rne(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  for (int d = 0; d < nv; d++) w->qfrc_passive[d] = -w->dof_damping[d] * w->qvel[d];
  real cacc[NB][6], cfrc[NB][6];
  for (int i = 0; i < 3; i++) { cacc[0][i] = 0; cacc[0][3 + i] = -(real)m->gravity[i]; }
  for (int b = 1; b < m->nbody; b++) {
    memcpy(cacc[b], cacc[m->body_parentid[b]], sizeof(cacc[b]));
    for (int k = 0; k < m->body_dofnum[b]; k++) {
      int d = m->body_dofadr[b] + k;
      for (int i = 0; i < 6; i++) cacc[b][i] += w->cdof_dot[d][i] * w->qvel[d];
    }
  }
  for (int b = 0; b < m->nbody; b++) {
    real f1[6], f2[6], f3[6];
    inert_mul(f1, w->cinert[b], cacc[b]);
    inert_mul(f2, w->cinert[b], w->cvel[b]);
    motion_cross_force(f3, w->cvel[b], f2);
    for (int i = 0; i < 6; i++) cfrc[b][i] = f1[i] + f3[i];
  }
  for (int b = m->nbody - 1; b > 0; b--) {
    int p = m->body_parentid[b];
    for (int i = 0; i < 6; i++) cfrc[p][i] += cfrc[b][i];
  }
  for (int d = 0; d < nv; d++) {
    real s = 0;
    const real* f = cfrc[m->dof_bodyid[d]];
    for (int i = 0; i < 6; i++) s += w->cdof[d][i] * f[i];
    w->qfrc_bias[d] = s;
  }
}

/* [upstream] forward.py::fwd_actuation + fwd_acceleration */
static void actuation_and_acceleration(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  for (int d = 0; d < nv; d++) w->qfrc_actuator[d] = 0;
  for (int u = 0; u < m->nu; u++) {
    int j = m->act_trnid[u], qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    real ctrl = w->ctrl[u];
    if (m->act_ctrllimited[u]) ctrl = clipr(ctrl, (real)m->act_ctrlrange[u][0], (real)m->act_ctrlrange[u][1]);
    real gear = (real)m->act_gear[u];
    real length = gear * w->qpos[qa], velocity = gear * w->qvel[d];
    real gain = (real)m->act_gainprm[u][0];
    real bias = (real)m->act_biasprm[u][0] + (real)m->act_biasprm[u][1] * length + (real)m->act_biasprm[u][2] * velocity;
    real force = gain * ctrl + bias;
    if (m->act_forcelimited[u]) force = clipr(force, (real)m->act_forcerange[u][0], (real)m->act_forcerange[u][1]);
    w->qfrc_actuator[d] += gear * force;
  }
  for (int d = 0; d < nv; d++) {
    int j = m->dof_jntid[d];
    if (m->jnt_actfrclimited[j])
      w->qfrc_actuator[d] = clipr(w->qfrc_actuator[d], (real)m->jnt_actfrcrange[j][0], (real)m->jnt_actfrcrange[j][1]);
  }
  for (int d = 0; d < nv; d++) {
    w->qfrc_smooth[d] = w->qfrc_passive[d] - w->qfrc_bias[d] + w->qfrc_actuator[d];
    w->qacc_smooth[d] = w->qfrc_smooth[d];
  }
  chol_solve(w->L, nv, w->qacc_smooth);
}

/* ---------------------------------------------------------------------- solver
 * [upstream] mjx/_src/solver.py (Newton, pyramidal cone, dense) */
typedef struct {
  real qacc[NV], Ma[NV], grad[NV], Mgrad[NV], search[NV];
  real gauss, cost, prev_cost;
} sctx;

static void mul_m(const work* w, const real* v, real* out) {
  int nv = w->m->nv;
  for (int i = 0; i < nv; i++) {
    real s = 0;
    for (int j = 0; j < nv; j++) s += w->M[i][j] * v[j];
    out[i] = s;
  }
}

/* _update_constraint: forces, active set, cost */
static void update_constraint(work* w, sctx* c) {
  int nv = w->m->nv, ne = w->ne, nf = w->nf;
  real cost = 0;
  for (int i = 0; i < nv; i++) w->qfrc_constraint[i] = 0;
  w->nefc_active = 0;
  for (int r = 0; r < w->nefc; r++) {
    real ja = w->Jaref[r], f;
    if (r < ne) {
      w->active[r] = 1;
      f = -w->D[r] * ja;
      cost += (real)0.5 * w->D[r] * ja * ja;
    } else if (r < ne + nf) {
      real fl = w->floss[r], rf = fl / w->D[r];
      if (ja <= -rf) { w->active[r] = 0; f = fl; cost += fl * ((real)-0.5 * rf - ja); }
      else if (ja >= rf) { w->active[r] = 0; f = -fl; cost += fl * ((real)-0.5 * rf + ja); }
      else { w->active[r] = 1; f = -w->D[r] * ja; cost += (real)0.5 * w->D[r] * ja * ja; }
    } else {
      w->active[r] = ja < 0;
      f = w->active[r] ? -w->D[r] * ja : 0;
      if (w->active[r]) cost += (real)0.5 * w->D[r] * ja * ja;
    }
    w->force[r] = f;
    if (w->active[r]) w->nefc_active++;
    if (f != 0)
      for (int i = 0; i < nv; i++) w->qfrc_constraint[i] += w->J[r][i] * f;
  }
  real gauss = 0;
  for (int i = 0; i < nv; i++) gauss += (c->Ma[i] - w->qfrc_smooth[i]) * (c->qacc[i] - w->qacc_smooth[i]);
  gauss *= (real)0.5;
  c->gauss = gauss;
  c->prev_cost = c->cost;
  c->cost = cost + gauss;
}

/* _update_gradient (Newton): grad, H = M + J^T diag(D*active) J, Mgrad = H^-1 grad */
static void update_gradient(work* w, sctx* c) {
  int nv = w->m->nv;
  static __thread real H[NV][NV], LH[NV][NV];
  for (int i = 0; i < nv; i++) c->grad[i] = c->Ma[i] - w->qfrc_smooth[i] - w->qfrc_constraint[i];
  for (int i = 0; i < nv; i++)
    for (int j = 0; j < nv; j++) H[i][j] = w->M[i][j];
  for (int r = 0; r < w->nefc; r++) {
    if (!w->active[r]) continue;
    real d = w->D[r];
    for (int i = 0; i < nv; i++) {
      real ji = w->J[r][i] * d;
      if (ji == 0) continue;
      for (int j = 0; j <= i; j++) H[i][j] += ji * w->J[r][j];
    }
  }
  for (int i = 0; i < nv; i++)
    for (int j = i + 1; j < nv; j++) H[i][j] = H[j][i];
  cholesky(LH, H, nv);
  for (int i = 0; i < nv; i++) c->Mgrad[i] = c->grad[i];
  chol_solve(LH, nv, c->Mgrad);
}

static void ctx_create(work* w, sctx* c, const real* qacc, int grad) {
  int nv = w->m->nv;
  for (int i = 0; i < nv; i++) c->qacc[i] = qacc[i];
  for (int r = 0; r < w->nefc; r++) {
    real s = 0;
    for (int i = 0; i < nv; i++) s += w->J[r][i] * qacc[i];
    w->Jaref[r] = s - w->aref[r];
  }
  mul_m(w, qacc, c->Ma);
  c->cost = INFINITY;
  c->prev_cost = 0;
  c->gauss = 0;
  update_constraint(w, c);
  if (grad) {
    update_gradient(w, c);
    for (int i = 0; i < nv; i++) c->search[i] = -c->Mgrad[i];
  }
}

typedef struct { real alpha, cost, deriv0, deriv1; } lspoint;

/* _LSPoint.create */
static lspoint ls_point(const work* w, real alpha, const real* quad_gauss) {
  int ne = w->ne, nf = w->nf;
  real q0 = quad_gauss[0], q1 = quad_gauss[1], q2 = quad_gauss[2];
  for (int r = 0; r < w->nefc; r++) {
    real x = w->Jaref[r] + alpha * w->Jv[r];
    if (r < ne) {
      q0 += w->quad[r][0]; q1 += w->quad[r][1]; q2 += w->quad[r][2];
    } else if (r < ne + nf) {
      real f = w->floss[r], rf = f / w->D[r];
      if (x <= -rf) { q0 += f * ((real)-0.5 * rf - w->Jaref[r]); q1 += -f * w->Jv[r]; }
      else if (x >= rf) { q0 += f * ((real)-0.5 * rf + w->Jaref[r]); q1 += f * w->Jv[r]; }
      else { q0 += w->quad[r][0]; q1 += w->quad[r][1]; q2 += w->quad[r][2]; }
    } else if (x < 0) {
      q0 += w->quad[r][0]; q1 += w->quad[r][1]; q2 += w->quad[r][2];
    }
  }
  lspoint p;
  p.alpha = alpha;
  p.cost = alpha * alpha * q2 + alpha * q1 + q0;
  p.deriv0 = 2 * alpha * q2 + q1;
  p.deriv1 = 2 * q2 + (q2 == 0 ? MJ_MINVAL : 0);
  return p;
}

/* _linesearch */
static void linesearch(work* w, sctx* c) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  real smag = 0, mv[NV];
  for (int i = 0; i < nv; i++) smag += c->search[i] * c->search[i];
  smag = R_SQRT(smag) * (real)m->meaninertia * (real)(nv > 1 ? nv : 1);
  real gtol = (real)m->tolerance * (real)m->ls_tolerance * smag;
  mul_m(w, c->search, mv);
  for (int r = 0; r < w->nefc; r++) {
    real s = 0;
    for (int i = 0; i < nv; i++) s += w->J[r][i] * c->search[i];
    w->Jv[r] = s;
    w->quad[r][0] = (real)0.5 * w->Jaref[r] * w->Jaref[r] * w->D[r];
    w->quad[r][1] = s * w->Jaref[r] * w->D[r];
    w->quad[r][2] = (real)0.5 * s * s * w->D[r];
  }
  real qg[3] = {c->gauss, 0, 0};
  for (int i = 0; i < nv; i++) {
    qg[1] += c->search[i] * c->Ma[i] - c->search[i] * w->qfrc_smooth[i];
    qg[2] += c->search[i] * mv[i];
  }
  qg[2] *= (real)0.5;
  lspoint p0 = ls_point(w, 0, qg);
  lspoint lo0 = ls_point(w, p0.alpha - p0.deriv0 / p0.deriv1, qg);
  int lesser = lo0.deriv0 < p0.deriv0;
  lspoint hi = lesser ? p0 : lo0;
  lspoint lo = lesser ? lo0 : p0;
  int swap = 1, it = 0;
  for (;;) {
    int done = it >= m->ls_iterations;
    done |= !swap;
    done |= (lo.deriv0 < 0) && (lo.deriv0 > -gtol);
    done |= (hi.deriv0 > 0) && (hi.deriv0 < gtol);
    if (done) break;
    lspoint lo_next = ls_point(w, lo.alpha - lo.deriv0 / lo.deriv1, qg);
    lspoint hi_next = ls_point(w, hi.alpha - hi.deriv0 / hi.deriv1, qg);
    lspoint mid = ls_point(w, (real)0.5 * (lo.alpha + hi.alpha), qg);
    int swap_lo_next = (lo.deriv0 > 0) || (lo.deriv0 < lo_next.deriv0);
    if (swap_lo_next) lo = lo_next;
    int swap_lo_mid = (mid.deriv0 < 0) && (lo.deriv0 < mid.deriv0);
    if (swap_lo_mid) lo = mid;
    int swap_hi_next = (hi.deriv0 < 0) || (hi.deriv0 > hi_next.deriv0);
    if (swap_hi_next) hi = hi_next;
    int swap_hi_mid = (mid.deriv0 > 0) && (hi.deriv0 > mid.deriv0);
    if (swap_hi_mid) hi = mid;
    swap = swap_lo_next | swap_lo_mid | swap_hi_next | swap_hi_mid;
    it++;
  }
  w->ls_total += it;
  int improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  real alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    for (int i = 0; i < nv; i++) { c->qacc[i] += c->search[i] * alpha; c->Ma[i] += mv[i] * alpha; }
    for (int r = 0; r < w->nefc; r++) w->Jaref[r] += w->Jv[r] * alpha;
  }
}

/* solver.solve */
static void solve(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  sctx c;
  w->solver_niter = 0;
  w->ls_total = 0;
  if (w->nefc == 0) {
    for (int i = 0; i < nv; i++) { w->qacc[i] = w->qacc_smooth[i]; w->qfrc_constraint[i] = 0; }
    return; /* forward(): "if d.efc_J.size == 0" — qacc_warmstart is left untouched */
  }
  /* warmstart: whichever of qacc_warmstart / qacc_smooth costs less */
  ctx_create(w, &c, w->warm, 0);
  real cw = c.cost;
  ctx_create(w, &c, w->qacc_smooth, 0);
  real cs = c.cost;
  ctx_create(w, &c, cw < cs ? w->warm : w->qacc_smooth, 1);
  real scale = (real)1 / ((real)m->meaninertia * (real)(nv > 1 ? nv : 1));
  for (;;) {
    real improvement = (c.prev_cost - c.cost) * scale;
    real g = 0;
    for (int i = 0; i < nv; i++) g += c.grad[i] * c.grad[i];
    real gradient = R_SQRT(g) * scale;
    int done = w->solver_niter >= m->iterations;
    done |= improvement < (real)m->tolerance;
    done |= gradient < (real)m->tolerance;
    if (done) break;
    linesearch(w, &c);
    update_constraint(w, &c);
    update_gradient(w, &c);
    for (int i = 0; i < nv; i++) c.search[i] = -c.Mgrad[i];
    w->solver_niter++;
  }
  for (int i = 0; i < nv; i++) { w->qacc[i] = c.qacc[i]; w->warm[i] = c.qacc[i]; }
}

/* ------------------------------------------------------------- forward / step
 * [upstream] forward.py::forward */
static void forward(work* w) {
  kinematics(w);
  com_pos(w);
  crb_and_factor(w);
  collision(w);
  make_constraint(w);
  com_vel(w);
  passive_and_rne(w);
  actuation_and_acceleration(w);
  solve(w);
}

/* [upstream] forward.py::implicit + _advance (integrator implicitfast: only the
 * dof-damping derivative survives for these models) */
static void implicit_advance(work* w) {
  const rsrx_model_blob* m = w->m;
  int nv = m->nv;
  real dt = (real)m->timestep;
  static __thread real MH[NV][NV], LH[NV][NV];
  real qacc[NV];
  for (int i = 0; i < nv; i++) {
    for (int j = 0; j < nv; j++) MH[i][j] = w->M[i][j];
    MH[i][i] += dt * w->dof_damping[i];
    qacc[i] = w->qfrc_smooth[i] + w->qfrc_constraint[i];
  }
  cholesky(LH, MH, nv);
  chol_solve(LH, nv, qacc);
  for (int i = 0; i < nv; i++) w->qvel[i] += qacc[i] * dt;
  for (int j = 0; j < m->njnt; j++) {
    int qa = m->jnt_qposadr[j], d = m->jnt_dofadr[j];
    if (m->jnt_type[j] == RSRX_JNT_FREE) {
      for (int i = 0; i < 3; i++) w->qpos[qa + i] += dt * w->qvel[d + i];
      /* math.quat_integrate */
      real v[3] = {w->qvel[d + 3], w->qvel[d + 4], w->qvel[d + 5]};
      real nrm = norm3(v);
      if (nrm > 0) for (int i = 0; i < 3; i++) v[i] /= nrm;
      real ang = dt * nrm, s = R_SIN(ang * (real)0.5), cc = R_COS(ang * (real)0.5);
      real qr[4] = {cc, v[0] * s, v[1] * s, v[2] * s}, q2[4];
      quat_mul(q2, w->qpos + qa + 3, qr);
      normalize_n(q2, 4);
      for (int i = 0; i < 4; i++) w->qpos[qa + 3 + i] = q2[i];
    } else {
      w->qpos[qa] += dt * w->qvel[d];
    }
  }
  w->time += dt;
}

static void step1(work* w) {
  forward(w);
  implicit_advance(w);
}

/* ------------------------------------------------------------- data <-> work */
static void load_model(work* w, const rsrx_model_blob* m, int dense) {
  w->m = m;
  w->dense = dense;
  for (int g = 0; g < m->ngeom; g++) for (int i = 0; i < 3; i++) w->geom_friction[g][i] = (real)m->geom_friction[g][i];
  for (int b = 0; b < m->nbody; b++) w->body_mass[b] = (real)m->body_mass[b];
  for (int d = 0; d < m->nv; d++) { w->dof_damping[d] = (real)m->dof_damping[d]; w->dof_frictionloss[d] = (real)m->dof_frictionloss[d]; }
}
static void load_data(work* w, const orc_data* d) {
  const rsrx_model_blob* m = w->m;
  for (int i = 0; i < m->nq; i++) w->qpos[i] = (real)d->qpos[i];
  for (int i = 0; i < m->nv; i++) { w->qvel[i] = (real)d->qvel[i]; w->warm[i] = (real)d->qacc_warmstart[i]; }
  for (int i = 0; i < m->nu; i++) w->ctrl[i] = (real)d->ctrl[i];
  w->time = (real)d->time;
}
static void store_data(const work* w, orc_data* d) {
  const rsrx_model_blob* m = w->m;
  for (int i = 0; i < m->nq; i++) d->qpos[i] = w->qpos[i];
  for (int i = 0; i < m->nv; i++) {
    d->qvel[i] = w->qvel[i]; d->qacc_warmstart[i] = w->warm[i]; d->qacc[i] = w->qacc[i];
    d->qacc_smooth[i] = w->qacc_smooth[i]; d->qfrc_constraint[i] = w->qfrc_constraint[i];
    d->qfrc_bias[i] = w->qfrc_bias[i]; d->qfrc_actuator[i] = w->qfrc_actuator[i];
  }
  for (int i = 0; i < m->nu; i++) d->ctrl[i] = w->ctrl[i];
  d->time = w->time;
  for (int b = 0; b < m->nbody; b++) {
    for (int i = 0; i < 3; i++) d->xpos[b][i] = w->xpos[b][i];
    for (int i = 0; i < 4; i++) d->xquat[b][i] = w->xquat[b][i];
  }
  for (int s = 0; s < m->nsite; s++) for (int i = 0; i < 3; i++) d->site_xpos[s][i] = w->sxpos[s][i];
  for (int g = 0; g < m->ngeom; g++) for (int i = 0; i < 3; i++) d->geom_xpos[g][i] = w->gxpos[g][i];
  d->ncon = w->ncon; d->ncon_active = w->ncon_active; d->nefc = w->nefc; d->nefc_active = w->nefc_active;
  d->solver_niter = w->solver_niter; d->ls_total = w->ls_total;
}

/* ============================================================ environment logic */
/* _get_obs: [ref] test/airbot.py:254-268, T_shape_env.py:223-234 */
static void get_obs(const rsrx_model_blob* m, const rsrx_env_cfg* cfg, const orc_env_state* s, const orc_data* d, double* obs_out) {
  (void)m;
  real obs[ORC_MAXOBS];
  int n = 0;
  for (int i = 0; i < 6; i++) obs[n++] = (real)d->qpos[cfg->joint_qadr[i]];
  if (cfg->env_kind == RSRX_ENV_T) {
    obs[n++] = (real)d->site_xpos[cfg->site_endpoint][2];
    for (int i = 0; i < 3; i++) obs[n++] = (real)s->target_pos[i] - (real)d->geom_xpos[cfg->geom_base][i];
    for (int i = 0; i < 3; i++) obs[n++] = (real)s->target2_pos[i] - (real)d->geom_xpos[cfg->geom_vertical][i];
    obs[n++] = (real)s->xita;
    for (int i = 0; i < 2; i++) obs[n++] = (real)s->new_pos[i] - (real)d->site_xpos[cfg->site_endpoint][i];
  } else {
    const double* cube = d->xpos[cfg->cube_body];
    const double* site = d->site_xpos[cfg->site_endpoint];
    for (int i = 0; i < 3; i++) obs[n++] = (real)site[i];
    for (int i = 0; i < 3; i++) obs[n++] = (real)s->target_pos[i];
    for (int i = 0; i < 3; i++) obs[n++] = (real)cube[i];
    for (int i = 0; i < 2; i++) obs[n++] = (real)s->new_pos[i];
    for (int i = 0; i < 3; i++) obs[n++] = (real)s->target_pos[i] - (real)cube[i];
    for (int i = 0; i < 3; i++) obs[n++] = (real)cube[i] - (real)site[i];
  }
  for (int i = 0; i < n; i++) obs_out[i] = obs[i];
  for (int i = n; i < ORC_MAXOBS; i++) obs_out[i] = 0;
}

/* reset: the caller samples qpos/qvel/ctrl exactly as the reference's reset does
 * ([ref] test/airbot.py:102-133) — this performs pipeline_init (= make_data +
 * mjx.forward with ctrl=0), data.replace(ctrl=...), info/metrics/obs init
 * ([ref] test/airbot.py:135-163, T_shape_env.py:111-137) and the wrappers' reset
 * (first_pipeline_state / first_obs / steps / truncation). */
int oracle_env_reset(const rsrx_model_blob* m, const rsrx_env_cfg* cfg, const double* qpos, const double* qvel,
                     const double* ctrl, orc_env_state* s, int dense) {
  work* w = (work*)calloc(1, sizeof(work));
  if (!w) return -1;
  load_model(w, m, dense);
  memset(s, 0, sizeof(*s));
  for (int i = 0; i < m->nq; i++) s->d.qpos[i] = qpos[i];
  for (int i = 0; i < m->nv; i++) s->d.qvel[i] = qvel[i];
  load_data(w, &s->d); /* ctrl = 0, warmstart = 0, time = 0 (make_data) */
  forward(w);
  store_data(w, &s->d);
  for (int i = 0; i < m->nu; i++) s->d.ctrl[i] = (real)ctrl[i];
  if (cfg->env_kind == RSRX_ENV_T) {
    s->new_pos[0] = (real)0.24739072; s->new_pos[1] = (real)-0.00496255;
    for (int i = 0; i < 3; i++) {
      s->target_pos[i] = s->d.geom_xpos[cfg->geom_target_base][i];
      s->target2_pos[i] = s->d.geom_xpos[cfg->geom_target_vertical][i];
      s->site_pos[i] = s->d.site_xpos[cfg->site_endpoint][i];
      s->obj_pos[i] = s->d.xpos[cfg->cube_body][i];
    }
    s->target_w = (real)(s->d.xquat[cfg->target_body][0]) * (real)10;
    s->xita = (real)0.2876;
  } else {
    s->new_pos[0] = (real)0.37342; s->new_pos[1] = (real)-0.07989;
    for (int i = 0; i < 3; i++) {
      s->target_pos[i] = s->d.xpos[cfg->target_body][i];
      s->site_pos[i] = s->d.site_xpos[cfg->site_endpoint][i];
      s->obj_pos[i] = s->d.xpos[cfg->cube_body][i];
    }
    s->last_action = 0;
  }
  get_obs(m, cfg, s, &s->d, s->obs);
  s->reward = 0; s->done = 0; s->truncation = 0; s->steps = 0;
  s->first = s->d;
  memcpy(s->first_obs, s->obs, sizeof(s->obs));
  free(w);
  return 0;
}

/* One wrapped env step: AutoResetWrapper(EpisodeWrapper(env)).step
 * [ref] env: test/airbot.py:165-252, cube_env.py:145-213, T_shape_env.py:139-221
 * [ref] wrappers: mujoco_playground/_src/wrapper.py:117-138 (AutoReset twin);
 * [upstream] brax/envs/wrappers/training.py EpisodeWrapper, AutoResetWrapper */
static void env_step_w(work* w, const rsrx_env_cfg* cfg, orc_env_state* s, const double* action_in) {
  const rsrx_model_blob* m = w->m;
  int kind = cfg->env_kind, nu = m->nu;
  /* episode_length <= 0: the bare (unwrapped) env, as rsr_pipeline.env_params_tuning steps it */
  const int wrapped = cfg->episode_length > 0;
  /* AutoReset pre: steps <- 0 where done; done <- 0 */
  if (wrapped && s->done != 0) s->steps = 0;
  s->done = 0;
  /* ---- pre-physics action shaping */
  orc_data* d0 = &s->d;
  real act[NU];
  for (int i = 0; i < nu; i++) act[i] = (real)d0->ctrl[i] + (real)cfg->action_scale[i] * (real)action_in[i];
  act[3] = -((real)1.57 + (real)d0->qpos[cfg->joint_qadr[1]] + (real)d0->qpos[cfg->joint_qadr[2]]);
  if (kind == RSRX_ENV_T) {
    real px = (real)d0->site_xpos[cfg->site_endpoint][0], py = (real)d0->site_xpos[cfg->site_endpoint][1];
    real dx = (real)d0->site_xpos[cfg->site_tail][0] - px, dy = (real)d0->site_xpos[cfg->site_tail][1] - py;
    real ang = R_ATAN2(dy, dx + (real)0.00001);
    act[4] = -ang + act[0] + (real)1.5708;
  } else {
    real cx = (real)d0->xpos[cfg->cube_body][0], cy = (real)d0->xpos[cfg->cube_body][1];
    real dx = (real)s->target_pos[0] - cx, dy = (real)s->target_pos[1] - cy;
    real ang = R_ATAN2(dy, dx + (real)0.00001);
    real a4 = -ang + act[0] + (real)1.5708;
    if (kind == RSRX_ENV_SF) {
      real df[3];
      for (int i = 0; i < 3; i++) df[i] = (real)s->target_pos[i] - (real)d0->xpos[cfg->cube_body][i];
      real dis0 = norm3(df);
      if (dis0 < (real)0.03) a4 = (real)s->last_action;
      act[4] = a4;
      s->last_action = act[4]; /* stored before the clip, [ref] test/airbot.py:184 */
    } else {
      act[4] = a4;
    }
  }
  for (int i = 0; i < nu; i++) act[i] = clipr(act[i], (real)m->act_ctrlrange[i][0], (real)m->act_ctrlrange[i][1]);
  /* ---- pipeline_step: n_frames x mjx.step with ctrl = action */
  load_data(w, d0);
  for (int i = 0; i < nu; i++) w->ctrl[i] = act[i];
  for (int f = 0; f < cfg->n_frames; f++) step1(w);
  store_data(w, &s->d);
  orc_data* d1 = &s->d;
  /* ---- post-physics */
  real reward, done;
  real site[3] = {(real)d1->site_xpos[cfg->site_endpoint][0], (real)d1->site_xpos[cfg->site_endpoint][1], (real)d1->site_xpos[cfg->site_endpoint][2]};
  real W = (real)cfg->siet_to_box_reward_weight;
  if (kind == RSRX_ENV_T) {
    real db[3], dv[3], box[3], tgt[3];
    for (int i = 0; i < 3; i++) {
      db[i] = (real)s->target_pos[i] - (real)d1->geom_xpos[cfg->geom_base][i];
      dv[i] = (real)s->target2_pos[i] - (real)d1->geom_xpos[cfg->geom_vertical][i];
      box[i] = (real)d1->geom_xpos[cfg->geom_vertical][i] - (real)d1->geom_xpos[cfg->geom_base][i];
      tgt[i] = (real)s->target2_pos[i] - (real)s->target_pos[i];
    }
    real disb = norm3(db), disv = norm3(dv);
    if (disb < (real)0.005) disb = 0;
    if (disv < (real)0.005) disv = 0;
    real prb = (real)1 / (1 + (real)10 * disb), prv = (real)1 / (1 + (real)10 * disv);
    real cosx = dot3(box, tgt) / (norm3(box) * norm3(tgt));
    real xita = R_ACOS(clipr(cosx, -1, 1));
    s->xita = xita;
    real pw_ = (real)1 / (1 + (real)6 * xita);
    real push = ((real)0.1515 * prb + (real)0.1515 * prv + (real)0.66 * pw_) * (real)cfg->push_reward_weight;
    real tail[3] = {(real)d1->site_xpos[cfg->site_tail][0], (real)d1->site_xpos[cfg->site_tail][1], 0};
    real old_new[2] = {(real)s->new_pos[0], (real)s->new_pos[1]};
    real site_z_reward = site[2] < (real)0.83 ? (real)1 : (real)0;
    real zdis = R_FABS(site[2] - (real)0.805);
    site_z_reward += (real)4 / (1 + 3 * zdis);
    real dx = (real)d1->site_xpos[cfg->site_target_tail][0] - tail[0];
    real dy = (real)d1->site_xpos[cfg->site_target_tail][1] - tail[1];
    real ang = R_ATAN2(dy, dx + (real)0.00001);
    real distance = R_SQRT(dx * dx + dy * dy) + (real)0.025;
    real y_ = distance * R_SIN(ang), x_ = distance * R_COS(ang);
    s->new_pos[0] = dx - x_ + tail[0];
    s->new_pos[1] = dy - y_ + tail[1];
    real e[2] = {site[0] - old_new[0], site[1] - old_new[1]};
    real sd = R_SQRT(e[0] * e[0] + e[1] * e[1]);
    sd = sd < (real)0.02 ? 0 : sd - (real)0.02;
    real s2c = (1 - R_TANH(5 * sd)) * W;
    real health = (real)cfg->healthy_reward * R_FABS((site[2] < (real)cfg->endpoint_min_z_pos ? (real)1 : (real)0) - 1);
    reward = push + s2c + health + site_z_reward;
    done = (real)d1->xpos[cfg->cube_body][2] < (real)0.6 ? 1 : 0;
    s->metrics[0] = push; s->metrics[1] = s2c; s->metrics[2] = health; s->metrics[4] = site_z_reward;
  } else {
    real cube[3] = {(real)d1->xpos[cfg->cube_body][0], (real)d1->xpos[cfg->cube_body][1], (real)d1->xpos[cfg->cube_body][2]};
    real df[3] = {(real)s->target_pos[0] - cube[0], (real)s->target_pos[1] - cube[1], (real)s->target_pos[2] - cube[2]};
    real dis = norm3(df);
    real thr = kind == RSRX_ENV_SF ? (real)0.003 : (real)0.005;
    if (dis < thr) dis = 0;
    real push = ((real)1 / (1 + 3 * dis)) * (real)cfg->push_reward_weight;
    real task_complete = (kind == RSRX_ENV_SF && dis < (real)0.003) ? (real)5 : (real)0;
    real old_new[2] = {(real)s->new_pos[0], (real)s->new_pos[1]};
    real site_z_reward = site[2] < (real)0.82 ? (real)1 : (real)0;
    real dx = (real)s->target_pos[0] - cube[0], dy = (real)s->target_pos[1] - cube[1];
    real ang = R_ATAN2(dy, dx + (real)0.00001);
    real distance = R_SQRT(dx * dx + dy * dy) + (real)0.04;
    real y_ = distance * R_SIN(ang), x_ = distance * R_COS(ang);
    s->new_pos[0] = dx - x_ + cube[0];
    s->new_pos[1] = dy - y_ + cube[1];
    real e[2] = {site[0] - old_new[0], site[1] - old_new[1]};
    real sd = R_SQRT(e[0] * e[0] + e[1] * e[1]);
    sd = sd < (real)0.042 ? 0 : sd - (real)0.042;
    real s2c = (1 - R_TANH(5 * sd)) * W;
    if (dis < (real)0.005) s2c = W;
    real health;
    if (kind == RSRX_ENV_SF) {
      real dn = 0;
      if (site[2] < (real)cfg->endpoint_min_z_pos) dn = 1;
      if (site[0] > (real)1.0) dn = 1;
      if (site[0] < (real)-0.6) dn = 1;
      if (site[1] > (real)0.3) dn = 1;
      if (site[1] < (real)-0.3) dn = 1;
      if (cube[2] < (real)0.6) dn = 1;
      health = (real)cfg->healthy_reward * R_FABS(dn - 1);
      reward = push + s2c + health + task_complete + site_z_reward;
      done = dis < (real)0.003 ? 1 : 0;
      s->metrics[0] = push; s->metrics[1] = 0; s->metrics[2] = s2c;
    } else {
      health = (real)cfg->healthy_reward * R_FABS((site[2] < (real)cfg->endpoint_min_z_pos ? (real)1 : (real)0) - 1);
      reward = push + s2c + health + site_z_reward;
      done = cube[2] < (real)0.6 ? 1 : 0;
      s->metrics[0] = push; s->metrics[2] = s2c;
    }
  }
  reward = clipr(reward, (real)-1e2, (real)1e2);
  get_obs(m, cfg, s, d1, s->obs);
  for (int i = 0; i < 3; i++) { s->site_pos[i] = site[i]; s->obj_pos[i] = d1->xpos[cfg->cube_body][i]; }
  s->reward = reward;
  /* ---- EpisodeWrapper */
  s->steps += cfg->action_repeat;
  if (wrapped && s->steps >= cfg->episode_length) { s->truncation = 1 - done; done = 1; }
  else s->truncation = 0;
  s->done = done;
  /* ---- AutoReset post: pipeline_state and obs only */
  if (wrapped && done != 0) {
    s->d = s->first;
    memcpy(s->obs, s->first_obs, sizeof(s->obs));
  }
}

int oracle_env_step(const rsrx_model_blob* m, const rsrx_env_cfg* cfg, orc_env_state* s, const double* action, int dense) {
  work* w = (work*)calloc(1, sizeof(work));
  if (!w) return -1;
  load_model(w, m, dense);
  env_step_w(w, cfg, s, action);
  free(w);
  return 0;
}

/* N envs x T steps; models[i*model_stride] is env i's (possibly randomised)
 * model; actions are [T][N][nu]; OpenMP over envs.  Returns 0. */
int oracle_rollout(const rsrx_model_blob* models, int model_stride, const rsrx_env_cfg* cfg, orc_env_state* states,
                   int N, const double* actions, int T, int dense, int nthreads) {
  int nu = models[0].nu;
#ifdef _OPENMP
  if (nthreads > 0) omp_set_num_threads(nthreads);
#endif
#pragma omp parallel
  {
    work* w = (work*)calloc(1, sizeof(work));
#pragma omp for schedule(dynamic, 4)
    for (int i = 0; i < N; i++) {
      load_model(w, &models[(size_t)i * model_stride], dense);
      for (int t = 0; t < T; t++) env_step_w(w, cfg, &states[i], actions + ((size_t)t * N + i) * nu);
    }
    free(w);
  }
  return 0;
}

/* ------------------------------------------------------ physics-only entry points */
int oracle_forward(const rsrx_model_blob* m, orc_data* d, int dense) {
  work* w = (work*)calloc(1, sizeof(work));
  if (!w) return -1;
  load_model(w, m, dense);
  load_data(w, d);
  forward(w);
  store_data(w, d);
  free(w);
  return 0;
}
int oracle_step(const rsrx_model_blob* m, orc_data* d, int nsteps, int dense) {
  work* w = (work*)calloc(1, sizeof(work));
  if (!w) return -1;
  load_model(w, m, dense);
  load_data(w, d);
  for (int i = 0; i < nsteps; i++) step1(w);
  store_data(w, d);
  free(w);
  return 0;
}

/* ---------------------------------------------------------- inspection (tests) */
/* forward() then dump internals: M[nv*nv], contacts, efc arrays (J[nefc*nv]...) */
int oracle_inspect(const rsrx_model_blob* m, orc_data* d, int dense, double* M, orc_contact* con, int* ncon, double* J,
                   double* D, double* aref, double* force, int* nefc, double* cdof, double* subtree_com) {
  work* w = (work*)calloc(1, sizeof(work));
  if (!w) return -1;
  load_model(w, m, dense);
  load_data(w, d);
  forward(w);
  store_data(w, d);
  int nv = m->nv;
  if (M) for (int i = 0; i < nv; i++) for (int j = 0; j < nv; j++) M[i * nv + j] = w->M[i][j];
  if (con) {
    for (int c = 0; c < w->ncon; c++) {
      con[c].dist = w->con[c].dist;
      for (int i = 0; i < 3; i++) con[c].pos[i] = w->con[c].pos[i];
      for (int i = 0; i < 9; i++) con[c].frame[i] = w->con[c].frame[i];
      for (int i = 0; i < 5; i++) { con[c].friction[i] = w->con[c].friction[i]; con[c].solimp[i] = w->con[c].solimp[i]; }
      for (int i = 0; i < 2; i++) con[c].solref[i] = w->con[c].solref[i];
      con[c].geom1 = w->con[c].g1; con[c].geom2 = w->con[c].g2;
    }
  }
  if (ncon) *ncon = w->ncon;
  if (J) for (int r = 0; r < w->nefc; r++) for (int i = 0; i < nv; i++) J[r * nv + i] = w->J[r][i];
  if (D) for (int r = 0; r < w->nefc; r++) D[r] = w->D[r];
  if (aref) for (int r = 0; r < w->nefc; r++) aref[r] = w->aref[r];
  if (force) for (int r = 0; r < w->nefc; r++) force[r] = w->force[r];
  if (nefc) *nefc = w->nefc;
  if (cdof) for (int i = 0; i < nv; i++) for (int k = 0; k < 6; k++) cdof[i * 6 + k] = w->cdof[i][k];
  if (subtree_com) for (int b = 0; b < m->nbody; b++) for (int k = 0; k < 3; k++) subtree_com[b * 3 + k] = w->subtree_com[b][k];
  free(w);
  return 0;
}

/* standalone narrow-phase for unit tests: mats are row-major 3x3 */
void oracle_box_box(const double* p1, const double* m1, const double* s1, const double* p2, const double* m2,
                    const double* s2, double* dist, double* pos, double* nrm) {
  real P1[3], M1[9], S1[3], P2[3], M2[9], S2[3], di[4], po[4][3], n[3];
  for (int i = 0; i < 3; i++) { P1[i] = (real)p1[i]; S1[i] = (real)s1[i]; P2[i] = (real)p2[i]; S2[i] = (real)s2[i]; }
  for (int i = 0; i < 9; i++) { M1[i] = (real)m1[i]; M2[i] = (real)m2[i]; }
  box_box(P1, M1, S1, P2, M2, S2, di, po, n);
  for (int c = 0; c < 4; c++) { dist[c] = di[c]; for (int i = 0; i < 3; i++) pos[c * 3 + i] = po[c][i]; }
  for (int i = 0; i < 3; i++) nrm[i] = n[i];
}
void oracle_plane_box(const double* pp, const double* pm, const double* bp, const double* bm, const double* size,
                      double* dist, double* pos, double* nrm) {
  real PP[3], PM[9], BP[3], BM[9], S[3], di[4], po[4][3], n[3];
  for (int i = 0; i < 3; i++) { PP[i] = (real)pp[i]; BP[i] = (real)bp[i]; S[i] = (real)size[i]; }
  for (int i = 0; i < 9; i++) { PM[i] = (real)pm[i]; BM[i] = (real)bm[i]; }
  plane_box(PP, PM, BP, BM, S, di, po, n);
  for (int c = 0; c < 4; c++) { dist[c] = di[c]; for (int i = 0; i < 3; i++) pos[c * 3 + i] = po[c][i]; }
  for (int i = 0; i < 3; i++) nrm[i] = n[i];
}

int oracle_sizeof_data(void) { return (int)sizeof(orc_data); }
int oracle_sizeof_env_state(void) { return (int)sizeof(orc_env_state); }
int oracle_sizeof_model(void) { return (int)sizeof(rsrx_model_blob); }
int oracle_sizeof_cfg(void) { return (int)sizeof(rsrx_env_cfg); }
int oracle_real_bytes(void) { return (int)sizeof(real); }
int oracle_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
