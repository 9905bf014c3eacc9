// rsrx_physics.cuh — one warp = one environment.  Every stage of mjx.step
// (mujoco-mjx 3.2.4 forward.py::step) as a warp-cooperative
// device function over the per-warp shared-memory arena (rsrx_device.cuh).
// Lanes map to bodies / dofs / geoms / geom pairs / constraint rows; tree
// recursions run level by level; reductions use xor-shuffles.
#pragma once
#include "rsrx_device.cuh"

namespace RSRX_NS {

#define RSRX_SYNC() __syncwarp()

// ------------------------------------------------------------------ kinematics
// smooth.py::kinematics — static bodies/geoms come precomputed from the host.
__device__ __noinline__ void kinematics(const DModel* __restrict__ dm, float* sm, int lane) {
  const int nbody = dm->nbody;
  if (lane < nbody && dm->body_static[lane]) {
    for (int i = 0; i < 3; i++) sm[ar::XPOS + lane * 3 + i] = dm->static_xpos[lane][i];
    for (int i = 0; i < 4; i++) sm[ar::XQUAT + lane * 4 + i] = dm->static_xquat[lane][i];
  }
  RSRX_SYNC();
  for (int lev = 1; lev < dm->nlevel; ++lev) {
    const int b = lane;
    if (b < nbody && dm->body_depth[b] == lev && !dm->body_static[b]) {
      const int p = dm->body_parentid[b];
      float pq[4] = {sm[ar::XQUAT + p * 4], sm[ar::XQUAT + p * 4 + 1], sm[ar::XQUAT + p * 4 + 2], sm[ar::XQUAT + p * 4 + 3]};
      float bp[3] = {dm->body_pos[b][0], dm->body_pos[b][1], dm->body_pos[b][2]};
      float bq[4] = {dm->body_quat[b][0], dm->body_quat[b][1], dm->body_quat[b][2], dm->body_quat[b][3]};
      float pos[3], quat[4], t[3];
      rotate(t, bp, pq);
      for (int i = 0; i < 3; i++) pos[i] = sm[ar::XPOS + p * 3 + i] + t[i];
      quat_mul(quat, pq, bq);
      const int jn = dm->body_jntnum[b], ja0 = dm->body_jntadr[b];
      for (int k = 0; k < jn; k++) {
        const int j = ja0 + k, qa = dm->jnt_qposadr[j];
        const int jt = dm->jnt_type[j];
        if (jt == RSRX_JNT_FREE) {
          for (int i = 0; i < 3; i++) {
            pos[i] = sm[ar::QPOS + qa + i];
            sm[ar::XANCHOR + j * 3 + i] = pos[i];
            sm[ar::XAXIS + j * 3 + i] = (i == 2) ? 1.f : 0.f;
          }
          for (int i = 0; i < 4; i++) quat[i] = sm[ar::QPOS + qa + 3 + i];
          normalize4(quat);
          for (int i = 0; i < 4; i++) sm[ar::QPOS + qa + 3 + i] = quat[i];
        } else {
          float jp[3] = {dm->jnt_pos[j][0], dm->jnt_pos[j][1], dm->jnt_pos[j][2]};
          float jax[3] = {dm->jnt_axis[j][0], dm->jnt_axis[j][1], dm->jnt_axis[j][2]};
          float anchor[3], axis[3];
          rotate(anchor, jp, quat);
          for (int i = 0; i < 3; i++) anchor[i] += pos[i];
          rotate(axis, jax, quat);
          for (int i = 0; i < 3; i++) { sm[ar::XANCHOR + j * 3 + i] = anchor[i]; sm[ar::XAXIS + j * 3 + i] = axis[i]; }
          const float dq = sm[ar::QPOS + qa] - dm->qpos0[qa];
          if (jt == RSRX_JNT_HINGE) {
            float s, c;
            sincosf(dq * 0.5f, &s, &c);
            float ql[4] = {c, jax[0] * s, jax[1] * s, jax[2] * s}, q2[4];
            quat_mul(q2, quat, ql);
            for (int i = 0; i < 4; i++) quat[i] = q2[i];
            rotate(t, jp, quat);
            for (int i = 0; i < 3; i++) pos[i] = anchor[i] - t[i];
          } else {
            for (int i = 0; i < 3; i++) pos[i] += axis[i] * dq;
          }
        }
      }
      for (int i = 0; i < 3; i++) sm[ar::XPOS + b * 3 + i] = pos[i];
      for (int i = 0; i < 4; i++) sm[ar::XQUAT + b * 4 + i] = quat[i];
    }
    RSRX_SYNC();
  }
  if (lane < dm->nsite) {
    const int s = lane, b = dm->site_bodyid[s];
    float q[4] = {sm[ar::XQUAT + b * 4], sm[ar::XQUAT + b * 4 + 1], sm[ar::XQUAT + b * 4 + 2], sm[ar::XQUAT + b * 4 + 3]};
    float sp[3] = {dm->site_pos[s][0], dm->site_pos[s][1], dm->site_pos[s][2]}, t[3];
    rotate(t, sp, q);
    for (int i = 0; i < 3; i++) sm[ar::SXPOS + s * 3 + i] = sm[ar::XPOS + b * 3 + i] + t[i];
  }
  RSRX_SYNC();
}

// world position of a body's inertial frame origin (xipos), recomputed on demand
__device__ __forceinline__ void body_xipos(const DModel* __restrict__ dm, const float* sm, int b, float* out) {
  float q[4] = {sm[ar::XQUAT + b * 4], sm[ar::XQUAT + b * 4 + 1], sm[ar::XQUAT + b * 4 + 2], sm[ar::XQUAT + b * 4 + 3]};
  float ip[3] = {dm->body_ipos[b][0], dm->body_ipos[b][1], dm->body_ipos[b][2]}, t[3];
  rotate(t, ip, q);
  for (int i = 0; i < 3; i++) out[i] = sm[ar::XPOS + b * 3 + i] + t[i];
}

// world pose of a geom, recomputed on demand (static geoms come precomputed from the host)
__device__ __noinline__ void geom_pose(const DModel* __restrict__ dm, const float* sm, int g, float* pos, float* mat) {
  if (dm->geom_static[g]) {
    for (int i = 0; i < 3; i++) pos[i] = dm->geom_static_xpos[g][i];
    if (mat) for (int i = 0; i < 9; i++) mat[i] = dm->geom_static_xmat[g][i];
    return;
  }
  const int b = dm->geom_bodyid[g];
  float q[4] = {sm[ar::XQUAT + b * 4], sm[ar::XQUAT + b * 4 + 1], sm[ar::XQUAT + b * 4 + 2], sm[ar::XQUAT + b * 4 + 3]};
  float gp[3] = {dm->geom_pos[g][0], dm->geom_pos[g][1], dm->geom_pos[g][2]}, t[3];
  rotate(t, gp, q);
  for (int i = 0; i < 3; i++) pos[i] = sm[ar::XPOS + b * 3 + i] + t[i];
  if (mat) {
    float gq[4] = {dm->geom_quat[g][0], dm->geom_quat[g][1], dm->geom_quat[g][2], dm->geom_quat[g][3]}, q2[4];
    quat_mul(q2, q, gq);
    quat_to_mat(mat, q2);
  }
}

// smooth.py::com_pos — subtree_com of tree roots, cinert, cdof
__device__ __noinline__ void com_pos(const DModel* __restrict__ dm, float* sm, int lane) {
  const int nbody = dm->nbody;
  const float* bmass_e = reinterpret_cast<const float* const*>(sm + ar::PTRS)[1];
  // a root's subtree is the contiguous body range [b, subtree_end)
  if (lane < nbody && lane > 0 && dm->body_rootid[lane] == lane) {
    const int b = lane;
    float px = 0.f, py = 0.f, pz = 0.f, ms = 0.f, xi[3];
#pragma unroll 1
    for (int c = dm->body_subtree_end[b] - 1; c >= b; --c) {  // leaves first, like the reverse tree scan
      const float mc = bmass_e ? bmass_e[c] : dm->body_mass[c];
      body_xipos(dm, sm, c, xi);
      px += xi[0] * mc; py += xi[1] * mc; pz += xi[2] * mc;
      ms += mc;
    }
    float* sc = sm + ar::SCOM + dm->body_treeid[b] * 3;
    if (ms < MJ_MINVAL) {
      body_xipos(dm, sm, b, xi);
      sc[0] = xi[0]; sc[1] = xi[1]; sc[2] = xi[2];
    } else {
      const float d = fmaxf(ms, MJ_MINVAL);
      sc[0] = px / d; sc[1] = py / d; sc[2] = pz / d;
    }
  }
  RSRX_SYNC();
  if (lane < nbody && lane > 0 && !dm->body_static[lane]) {
    const int b = lane;
    const float* sc = sm + ar::SCOM + dm->body_treeid[b] * 3;
    float xi[3];
    body_xipos(dm, sm, b, xi);
    float off[3] = {xi[0] - sc[0], xi[1] - sc[1], xi[2] - sc[2]};
    const float ms = bmass_e ? bmass_e[b] : dm->body_mass[b];
    float q[4] = {sm[ar::XQUAT + b * 4], sm[ar::XQUAT + b * 4 + 1], sm[ar::XQUAT + b * 4 + 2], sm[ar::XQUAT + b * 4 + 3]};
    float iq[4] = {dm->body_iquat[b][0], dm->body_iquat[b][1], dm->body_iquat[b][2], dm->body_iquat[b][3]};
    float q2[4], R[9];
    quat_mul(q2, q, iq);
    quat_to_mat(R, q2);  // ximat
    const float I0 = dm->body_inertia[b][0], I1 = dm->body_inertia[b][1], I2 = dm->body_inertia[b][2];
    float in[6];  // 00 11 22 01 02 12
    const int rr[6] = {0, 1, 2, 0, 0, 1}, cc[6] = {0, 1, 2, 1, 2, 2};
    float h[3][3];
    for (int r3 = 0; r3 < 3; r3++) {
      float e[3] = {0.f, 0.f, 0.f};
      e[r3] = -1.f;
      cross3(h[r3], off, e);
    }
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const int r3 = rr[k], c3 = cc[k];
      float s = R[r3 * 3] * I0 * R[c3 * 3] + R[r3 * 3 + 1] * I1 * R[c3 * 3 + 1] + R[r3 * 3 + 2] * I2 * R[c3 * 3 + 2];
      in[k] = s + dot3(h[r3], h[c3]) * ms;
    }
    float* ci = sm + ar::CINERT + b * 10;
    for (int k = 0; k < 6; k++) ci[k] = in[k];
    ci[6] = off[0] * ms; ci[7] = off[1] * ms; ci[8] = off[2] * ms; ci[9] = ms;
  }
  if (lane < dm->nv) {
    const int d = lane, j = dm->dof_jntid[d], b = dm->jnt_bodyid[j], k = d - dm->jnt_dofadr[j];
    const int jt = dm->jnt_type[j];
    const float* sc = sm + ar::SCOM + dm->body_treeid[b] * 3;
    float off[3] = {sc[0] - sm[ar::XANCHOR + j * 3], sc[1] - sm[ar::XANCHOR + j * 3 + 1], sc[2] - sm[ar::XANCHOR + j * 3 + 2]};
    float* cd = sm + ar::CDOF + d * 6;
    if (jt == RSRX_JNT_FREE) {
      if (k < 3) {
        for (int i = 0; i < 6; i++) cd[i] = (i == 3 + k) ? 1.f : 0.f;
      } else {
        const int a = k - 3;
        float q[4] = {sm[ar::XQUAT + b * 4], sm[ar::XQUAT + b * 4 + 1], sm[ar::XQUAT + b * 4 + 2], sm[ar::XQUAT + b * 4 + 3]}, m[9];
        quat_to_mat(m, q);
        float ax[3] = {m[a], m[3 + a], m[6 + a]}, c[3];
        cross3(c, ax, off);
        cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2]; cd[3] = c[0]; cd[4] = c[1]; cd[5] = c[2];
      }
    } else if (jt == RSRX_JNT_SLIDE) {
      cd[0] = cd[1] = cd[2] = 0.f;
      for (int i = 0; i < 3; i++) cd[3 + i] = sm[ar::XAXIS + j * 3 + i];
    } else {
      float ax[3] = {sm[ar::XAXIS + j * 3], sm[ar::XAXIS + j * 3 + 1], sm[ar::XAXIS + j * 3 + 2]}, c[3];
      cross3(c, ax, off);
      cd[0] = ax[0]; cd[1] = ax[1]; cd[2] = ax[2]; cd[3] = c[0]; cd[4] = c[1]; cd[5] = c[2];
    }
  }
  RSRX_SYNC();
}

// Cholesky factorisation / solves of an SPD nv x nv matrix A held in the
// BLOCK-PERMUTED dof order (DModel::pos_of_dof: dofs that can ever be coupled —
// same kinematic tree or a collision pair between their trees — are contiguous;
// everything outside the diagonal blocks is structurally zero and never touched).
// A's lower triangle (packed by rows) is overwritten by its factor; 1/L_kk goes to
// ar::V_RDIAG.  COMPACT code on purpose (the kernel is instruction-fetch bound, see
// profiles/): rolled right-looking factorisation; lane p owns row p, which lives in
// shared memory (row starts 0, 1, 3, 6, ...: distinct banks for nv <= 20) and is touched by that lane
// only; whatever crosses lanes (pivot, column k) travels by shuffle, so there are
// no barriers in the dependency chain; the pivot column is scaled by one rsqrt
// (no IEEE sqrt / division subroutines).
__device__ __noinline__ void warp_chol_factor(const DModel* __restrict__ dm, float* sm, int lane, bool tree_blocks) {
  constexpr unsigned FULL = 0xffffffffu;
  float* A = sm + ar::HH;
  const int n = dm->nv;
  const bool own = lane < n;
  const int p = own ? lane : n - 1;  // surplus lanes shadow the last row (reads only)
  float* row = A + tri(p);
  // tree_blocks: no contact couples two kinematic trees right now (always true for M itself), so the blocks are
  // the trees (8 | 6 | 6 instead of 14 | 6 for the cube model).  The diagonal blocks are independent, so they are
  // factored SIDE BY SIDE: in round kk every block eliminates its own kk-th column (pivot lane pstart + kk), and the
  // dependency chain is as long as the largest block, not nv.  Per element the arithmetic is the sequential one.
  const int pstart = tree_blocks ? dm->tblk_start[p] : dm->blk_start[p];
  const int pend = tree_blocks ? dm->tblk_end[p] : dm->blk_end[p];
  const int rounds = tree_blocks ? dm->tblk_max : dm->blk_max;
  float rdiag = 1.f;
#pragma unroll 1
  for (int kk = 0; kk < rounds; ++kk) {
    const int k = pstart + kk;
    const bool act = own && k <= pend;  // my block still has a column kk
    const int kc = act ? k : p;
    const float aik = row[kc];
    const float akk = __shfl_sync(FULL, aik, kc);
    const float rd = rsqrtf(akk > MJ_MINVAL ? akk : MJ_MINVAL);
    const bool in = act && lane > k;
    const float l = in ? aik * rd : 0.f;
    if (in) row[k] = l;
    if (act && lane == k) { row[k] = akk * rd; rdiag = rd; }
#pragma unroll 2
    for (int jj = kk + 1; jj < rounds; ++jj) {
      const int j = pstart + jj;
      const float lj = __shfl_sync(FULL, l, j <= pend ? j : p);
      if (in && j <= lane) row[j] -= l * lj;
    }
  }
  if (own) sm[ar::V_RDIAG + lane] = rdiag;
  RSRX_SYNC();
}

// x <- (L L^T)^-1 x with the factor left in ar::HH by warp_chol_factor; x is an
// nv-vector in dof order.  Blocks side by side, as in the factorisation.
__device__ __noinline__ void warp_chol_solve(const DModel* __restrict__ dm, float* sm, float* x, int lane, bool tree_blocks) {
  constexpr unsigned FULL = 0xffffffffu;
  const float* A = sm + ar::HH;
  const int n = dm->nv;
  const bool own = lane < n;
  const int p = own ? lane : n - 1;
  const int pstart = tree_blocks ? dm->tblk_start[p] : dm->blk_start[p], pend = tree_blocks ? dm->tblk_end[p] : dm->blk_end[p];
  const int rounds = tree_blocks ? dm->tblk_max : dm->blk_max;
  const int dof = dm->dof_of_pos[p];
  const float rdiag = sm[ar::V_RDIAG + p];
  const float* row = A + tri(p);
  float xi = x[dof];
#pragma unroll 1
  for (int kk = 0; kk < rounds; ++kk) {  // forward: L y = b (row-oriented, own row)
    const int k = pstart + kk;
    const bool act = own && k <= pend;
    const float yk = __shfl_sync(FULL, xi * rdiag, act ? k : p);
    if (act && lane == k) xi = yk;
    else if (act && lane > k) xi -= row[k] * yk;
  }
#pragma unroll 1
  for (int kk = rounds - 1; kk >= 0; --kk) {  // backward: L^T x = y (column-oriented: row k is contiguous over lanes)
    const int k = pstart + kk;
    const bool act = own && k <= pend;
    const float xk = __shfl_sync(FULL, xi * rdiag, act ? k : p);
    if (act && lane == k) xi = xk;
    else if (act && lane < k) xi -= A[tri(k) + p] * xk;
  }
  if (own) x[dof] = xi;
  RSRX_SYNC();
}

// HH <- M (+ dt * diag(damp) when damp != nullptr), written in the block-permuted order
__device__ __noinline__ void copy_M_permuted(const DModel* __restrict__ dm, float* sm, int lane, const float* damp, float dt) {
#pragma unroll 1
  for (int e = lane; e < dm->ntri; e += 32) {
    const int dst = dm->tri_dst[e];
    float v = sm[ar::MM + e];  // packed lower triangle: entry e = (tri_i, tri_j)
    if (damp && dm->tri_i[e] == dm->tri_j[e]) v += dt * damp[dm->tri_i[e]];
    sm[ar::HH + dst] = v;
  }
  RSRX_SYNC();
}

// smooth.py::crb + support.make_m (dense) + factor_m
__device__ __noinline__ void crb_and_factor(const DModel* __restrict__ dm, float* sm, int lane) {
  const int nv = dm->nv, nbody = dm->nbody;
  if (lane < nbody && lane > 0 && !dm->body_static[lane]) {
    const int b = lane;
    float acc[10];
    for (int i = 0; i < 10; i++) acc[i] = 0.f;
    for (int c = dm->body_subtree_end[b] - 1; c >= b; --c)
      for (int i = 0; i < 10; i++) acc[i] += sm[ar::CINERT + c * 10 + i];
    for (int i = 0; i < 10; i++) sm[ar::CRB + b * 10 + i] = acc[i];
  }
  for (int e = lane; e < dm->ntri; e += 32) sm[ar::MM + e] = 0.f;
  RSRX_SYNC();
  if (lane < nv) {
    float f[6];
    inert_mul(f, sm + ar::CRB + dm->dof_bodyid[lane] * 10, sm + ar::CDOF + lane * 6);
    for (int k = 0; k < 6; k++) sm[ar::CDOFDOT + lane * 6 + k] = f[k];
  }
  RSRX_SYNC();
#pragma unroll 1
  for (int e = lane; e < dm->nment; e += 32) {
    const int i = dm->ment_i[e], j = dm->ment_j[e];  // j is an ancestor dof of i (j <= i)
    float s = 0.f;
    for (int k = 0; k < 6; k++) s += sm[ar::CDOF + j * 6 + k] * sm[ar::CDOFDOT + i * 6 + k];
    if (i == j) s += dm->dof_armature[i];
    sm[ar::MM + ((i * (i + 1)) >> 1) + j] = s;
  }
  RSRX_SYNC();
}

// ------------------------------------------------------------------- collision
// The narrow phase is WARP-COOPERATIVE: after the bounding cull one or two pairs survive per env, so instead of one
// lane per pair (a single lane walking 15 SAT axes through local memory) the whole warp works on one pair: one lane
// per SAT axis / polygon vertex / contact slot, arg-min/max by redux + ballot (first index wins, like the sequential
// scans of the oracle), everything in registers.
constexpr unsigned FULLM = 0xffffffffu;
// order-preserving float -> int (NaN lowest; -0 == +0)
__device__ __forceinline__ int fkey(float f) {
  f += 0.f;
  const int b = __float_as_int(f);
  return f != f ? INT_MIN : (b ^ ((b >> 31) & 0x7fffffff));
}
// A half warp (16 lanes) works on one geom pair; the two halves run side by side on two pairs.
struct Half {
  unsigned mask;  // member mask of this half
  int shift;      // 0 or 16
  int l;          // lane within the half
  __device__ __forceinline__ unsigned ballot(bool p) const { return (__ballot_sync(mask, p) >> shift) & 0xffffu; }
  __device__ __forceinline__ float shfl(float v, int src) const { return __shfl_sync(mask, v, src, 16); }
  __device__ __forceinline__ int shfl(int v, int src) const { return __shfl_sync(mask, v, src, 16); }
  __device__ __forceinline__ void shfl3(float* o, const float* v, int src) const {
    o[0] = shfl(v[0], src); o[1] = shfl(v[1], src); o[2] = shfl(v[2], src);
  }
  // first lane (of the half) holding the largest / smallest v among the valid lanes (0 if there is none)
  __device__ __forceinline__ int argmax_first(float v, bool valid) const {
    const int k = valid ? fkey(v) : INT_MIN;
    const int m = __reduce_max_sync(mask, k);
    return __ffs(ballot(k == m)) - 1;
  }
  __device__ __forceinline__ int argmin_first(float v, bool valid) const {
    const int k = (valid && v == v) ? fkey(v) : INT_MAX;
    const int m = __reduce_min_sync(mask, k);
    return __ffs(ballot(k == m)) - 1;
  }
};
__device__ __forceinline__ float sel3(const float* v, int k) { return k == 0 ? v[0] : (k == 1 ? v[1] : v[2]); }
// column k of a row-major 3x3 / unit vector k, k a run-time index (selects, no local memory)
__device__ __forceinline__ void col3(float* o, const float* R, int k) {
  o[0] = k == 0 ? R[0] : (k == 1 ? R[1] : R[2]);
  o[1] = k == 0 ? R[3] : (k == 1 ? R[4] : R[5]);
  o[2] = k == 0 ? R[6] : (k == 1 ? R[7] : R[8]);
}
__device__ __forceinline__ void unit3(float* o, int k) { o[0] = k == 0 ? 1.f : 0.f; o[1] = k == 1 ? 1.f : 0.f; o[2] = k == 2 ? 1.f : 0.f; }

// collision_convex.py::_manifold_points with deterministic tie-breaking (see oracle/rsr_oracle.c and DESIGN.md: a
// taken vertex scores like a masked one).  Lane i < n holds vertex i (P, mask); idx[] comes back in every lane.
__device__ void manifold_points(const float* P, bool mask, int n, const float* nrm, const Half& hw, int idx[4]) {
  const bool has = hw.l < n;
  float dmk = mask ? 0.f : -1e6f;
  const int a = hw.argmax_first(dmk, has);
  if (hw.l == a) dmk = -1e6f;
  float A[3], B[3], C[3];
  hw.shfl3(A, P, a);
  const float ap[3] = {A[0] - P[0], A[1] - P[1], A[2] - P[2]};
  const int b = hw.argmax_first(dot3(ap, ap) + dmk, has);
  if (hw.l == b) dmk = -1e6f;
  hw.shfl3(B, P, b);
  float ab[3];
  const float amb[3] = {A[0] - B[0], A[1] - B[1], A[2] - B[2]};
  cross3(ab, nrm, amb);
  const int c = hw.argmax_first(fabsf(dot3(ap, ab)) + dmk, has);
  if (hw.l == c) dmk = -1e6f;
  hw.shfl3(C, P, c);
  float ac[3], bc[3];
  const float amc[3] = {A[0] - C[0], A[1] - C[1], A[2] - C[2]};
  const float bmc[3] = {B[0] - C[0], B[1] - C[1], B[2] - C[2]};
  cross3(ac, nrm, amc);
  cross3(bc, nrm, bmc);
  const float bp[3] = {B[0] - P[0], B[1] - P[1], B[2] - P[2]};
  const float v1 = fabsf(dot3(bp, bc)) + dmk, v2 = fabsf(dot3(ap, ac)) + dmk;
  const int d1 = hw.argmax_first(v1, has), d2 = hw.argmax_first(v2, has);
  const float best1 = hw.shfl(v1, d1), best2 = hw.shfl(v2, d2);
  idx[0] = a; idx[1] = b; idx[2] = c; idx[3] = best2 > best1 ? d2 : d1;
}

// Result of a narrow phase: lane c < 4 of the half holds contact slot c (dist >= 0: none); nrm is uniform.
struct Manifold { float dist, pos[3], nrm[3]; };

// collision_convex.py::plane_convex on the 8 box vertices (lane k < 8 of the half: vertex k)
__device__ void plane_box(const float* ppos, const float* pmat, const float* bpos, const float* bmat, const float* size,
                          const Half& hw, Manifold* out) {
  float n[3], pp[3], d[3];
  const float pn[3] = {pmat[2], pmat[5], pmat[8]};
  for (int i = 0; i < 3; i++) d[i] = ppos[i] - bpos[i];
  matT_vec(pp, bmat, d);
  matT_vec(n, bmat, pn);
  const float vert[3] = {(hw.l & 1) ? size[0] : -size[0], (hw.l & 2) ? size[1] : -size[1], (hw.l & 4) ? size[2] : -size[2]};
  const float tv[3] = {pp[0] - vert[0], pp[1] - vert[1], pp[2] - vert[2]};
  const float support = dot3(tv, n);
  const float smax = hw.shfl(support, hw.argmax_first(support, hw.l < 8));
  float thr = smax - 1e-3f;
  if (thr < 0.f) thr = 0.f;
  int idx[4];
  manifold_points(vert, support > thr, 8, n, hw, idx);
  const int c = hw.l & 3;
  const int k = c == 0 ? idx[0] : c == 1 ? idx[1] : c == 2 ? idx[2] : idx[3];
  bool unique = true;
  if (c > 0 && idx[0] == k) unique = false;
  if (c > 1 && idx[1] == k) unique = false;
  if (c > 2 && idx[2] == k) unique = false;
  float vk[3], wp[3];
  hw.shfl3(vk, vert, k);
  const float sk = hw.shfl(support, k);
  mat_vec(wp, bmat, vk);
  out->dist = unique ? -sk : 1.f;
  for (int i = 0; i < 3; i++) { out->pos[i] = bpos[i] + wp[i] - 0.5f * out->dist * pn[i]; out->nrm[i] = pn[i]; }
}

// SAT + face clipping / edge-edge; same rules as the oracle.  Works in box 2's frame, where
// box 2's axes are the unit vectors and box 1's are the columns of R.  scratch: 24 floats of shared memory.
__device__ void box_box(const float* p1, const float* m1, const float* s1, const float* p2, const float* m2,
                        const float* s2, float* scratch, const Half& hw, Manifold* out) {
  float R[9], t[3], d[3];
  for (int i = 0; i < 3; i++) d[i] = p1[i] - p2[i];
  matT_vec(t, m2, d);
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) R[r * 3 + c] = m2[r] * m1[c] + m2[3 + r] * m1[3 + c] + m2[6 + r] * m1[6 + c];
  // --- 15 separating-axis candidates, one per lane
  const int k = hw.l < 15 ? hw.l : 14;
  float ax[3];
  bool degenerate = false;
  if (k < 3) col3(ax, R, k);
  else if (k < 6) unit3(ax, k - 3);
  else {
    float ea[3], eb[3];
    col3(ea, R, (k - 6) % 3);
    unit3(eb, (k - 6) / 3);
    cross3(ax, ea, eb);
    degenerate = dot3(ax, ax) < 1e-6f;
    normalize3(ax);
  }
  float ov, sg;
  {
    const float a0[3] = {R[0], R[3], R[6]}, a1[3] = {R[1], R[4], R[7]}, a2[3] = {R[2], R[5], R[8]};
    const float ca = dot3(t, ax);
    const float ra = s1[0] * fabsf(dot3(a0, ax)) + s1[1] * fabsf(dot3(a1, ax)) + s1[2] * fabsf(dot3(a2, ax));
    const float rb = s2[0] * fabsf(ax[0]) + s2[1] * fabsf(ax[1]) + s2[2] * fabsf(ax[2]);
    const float dist1 = (ca + ra) - (-rb);
    const float dist2 = rb - (ca - ra);
    sg = dist1 > dist2 ? -1.f : 1.f;
    ov = dist1 < dist2 ? dist1 : dist2;
    if (degenerate) ov = 1e6f;
  }
  const int bf = hw.argmin_first(ov, hw.l < 6), be = hw.argmin_first(ov, hw.l >= 6 && hw.l < 15);
  const float bestf = hw.shfl(ov, bf), beste = hw.shfl(ov, be);
  const bool is_edge = beste * 1.05f < bestf;
  const int best = is_edge ? be : bf;
  const float bov = is_edge ? beste : bestf;
  const float sgb = hw.shfl(sg, best);
  float n[3];
  hw.shfl3(n, ax, best);
  n[0] *= sgb; n[1] *= sgb; n[2] *= sgb;
  float dist = 1.f, lp[3] = {0.f, 0.f, 0.f};
  if (bov < 0.f) {
    // separated
  } else if (!is_edge) {
    const bool refA = best < 3;
    const int r = refA ? best : best - 3;
    // axP / axQ: axes of the reference / incident box
    float hP[3], hQ[3], cP[3], cQ[3], nref[3];
    for (int i = 0; i < 3; i++) {
      hP[i] = refA ? s1[i] : s2[i]; hQ[i] = refA ? s2[i] : s1[i];
      cP[i] = refA ? t[i] : 0.f; cQ[i] = refA ? 0.f : t[i]; nref[i] = refA ? n[i] : -n[i];
    }
    int q = 0;
    float bestd = -1.f;
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      float aq[3];
      if (refA) unit3(aq, kk); else col3(aq, R, kk);
      const float v = fabsf(dot3(aq, nref));
      if (v > bestd) { bestd = v; q = kk; }
    }
    const int u = (q + 1) % 3, v = (q + 2) % 3;
    float aq[3], au[3], av[3];
    if (refA) { unit3(aq, q); unit3(au, u); unit3(av, v); } else { col3(aq, R, q); col3(au, R, u); col3(av, R, v); }
    const float sq = dot3(aq, nref) > 0.f ? -1.f : 1.f;
    const float su = (hw.l == 0 || hw.l == 3) ? 1.f : -1.f, sv = hw.l < 2 ? 1.f : -1.f;
    const float hq = sel3(hQ, q), hu = sel3(hQ, u), hv = sel3(hQ, v);
    float P[3];
    for (int i = 0; i < 3; i++) P[i] = cQ[i] + sq * hq * aq[i] + su * hu * au[i] + sv * hv * av[i];
    int np = 4;
    const int pu = (r + 1) % 3, pv = (r + 2) % 3;
#pragma unroll 1
    for (int s = 0; s < 4 && np > 0; s++) {
      const int sa = s < 2 ? pu : pv;
      const float sgn = (s & 1) ? -1.f : 1.f;
      float pn[3];
      if (refA) col3(pn, R, sa); else unit3(pn, sa);
      pn[0] *= sgn; pn[1] *= sgn; pn[2] *= sgn;
      const float h = sel3(hP, sa) + dot3(cP, pn);
      // Sutherland-Hodgman against pn . x <= h: lane i owns edge (i, i + 1)
      const bool has = hw.l < np;
      const int nxt = hw.l + 1 < np ? hw.l + 1 : 0;
      const float da = dot3(P, pn) - h;
      float Q[3];
      hw.shfl3(Q, P, nxt);
      const float db = hw.shfl(da, nxt);
      const bool keep = has && da <= 0.f;
      const bool cut = has && ((da < 0.f && db > 0.f) || (da > 0.f && db < 0.f));
      const unsigned mk = hw.ballot(keep), mc = hw.ballot(cut), lt = (1u << hw.l) - 1u;
      const int off = __popc(mk & lt) + __popc(mc & lt);
      if (keep && off < 8) { scratch[off * 3] = P[0]; scratch[off * 3 + 1] = P[1]; scratch[off * 3 + 2] = P[2]; }
      if (cut && off + (keep ? 1 : 0) < 8) {
        const float tt = da / (da - db);
        float* o = scratch + (off + (keep ? 1 : 0)) * 3;
        o[0] = P[0] + tt * (Q[0] - P[0]); o[1] = P[1] + tt * (Q[1] - P[1]); o[2] = P[2] + tt * (Q[2] - P[2]);
      }
      __syncwarp(hw.mask);
      np = __popc(mk) + __popc(mc);
      if (np > 8) np = 8;
      if (hw.l < np) { P[0] = scratch[hw.l * 3]; P[1] = scratch[hw.l * 3 + 1]; P[2] = scratch[hw.l * 3 + 2]; }
      __syncwarp(hw.mask);
    }
    if (np > 0) {
      const float rel[3] = {P[0] - cP[0], P[1] - cP[1], P[2] - cP[2]};
      const float depth = sel3(hP, r) - dot3(rel, nref);
      const bool mask = depth > 0.f;
      const float ref[3] = {P[0] + depth * nref[0], P[1] + depth * nref[1], P[2] + depth * nref[2]};
      int idx[4];
      manifold_points(ref, mask, np, nref, hw, idx);
      const int c = hw.l & 3;
      const int kx = c == 0 ? idx[0] : c == 1 ? idx[1] : c == 2 ? idx[2] : idx[3];
      bool unique = true;
      if (c > 0 && idx[0] == kx) unique = false;
      if (c > 1 && idx[1] == kx) unique = false;
      if (c > 2 && idx[2] == kx) unique = false;
      float Pk[3];
      hw.shfl3(Pk, P, kx);
      const float dk = hw.shfl(depth, kx);
      const bool mk2 = hw.shfl(mask ? 1 : 0, kx) != 0;
      if (unique && mk2) {
        dist = -dk;
        for (int i = 0; i < 3; i++) lp[i] = Pk[i] + 0.5f * dk * nref[i];
      }
    }
  } else {
    // edge-edge: closest points of the two supporting edges (uniform arithmetic, reported in slot 0)
    const int ia = (best - 6) % 3, jb = (best - 6) / 3;
    float ea[3], eb[3];
    for (int i = 0; i < 3; i++) { ea[i] = t[i]; eb[i] = 0.f; }
#pragma unroll
    for (int kk = 0; kk < 3; kk++) {
      float aA[3], aB[3];
      col3(aA, R, kk);
      unit3(aB, kk);
      if (kk != ia) {
        const float sA = dot3(aA, n) >= 0.f ? 1.f : -1.f;
        for (int i = 0; i < 3; i++) ea[i] += sA * s1[kk] * aA[i];
      }
      if (kk != jb) {
        const float sB = dot3(aB, n) >= 0.f ? 1.f : -1.f;
        for (int i = 0; i < 3; i++) eb[i] -= sB * s2[kk] * aB[i];
      }
    }
    float ua[3], ub[3];
    col3(ua, R, ia);
    unit3(ub, jb);
    const float w0[3] = {ea[0] - eb[0], ea[1] - eb[1], ea[2] - eb[2]};
    const float bb = dot3(ua, ub), dd = dot3(ua, w0), ee = dot3(ub, w0);
    const float den = 1.f - bb * bb;
    float sa = den > 1e-12f ? (bb * ee - dd) / den : 0.f;
    float sb = den > 1e-12f ? (ee - bb * dd) / den : 0.f;
    const float h1 = sel3(s1, ia), h2 = sel3(s2, jb);
    sa = clipf(sa, -h1, h1);
    sb = clipf(sb, -h2, h2);
    float pa[3], pb[3], df[3];
    for (int i = 0; i < 3; i++) { pa[i] = ea[i] + sa * ua[i]; pb[i] = eb[i] + sb * ub[i]; df[i] = pb[i] - pa[i]; }
    if ((hw.l & 3) == 0) {
      dist = dot3(df, n);
      for (int i = 0; i < 3; i++) lp[i] = 0.5f * (pa[i] + pb[i]);
    }
  }
  float wp[3];
  mat_vec(wp, m2, lp);
  out->dist = dist;
  for (int i = 0; i < 3; i++) out->pos[i] = p2[i] + wp[i];
  mat_vec(out->nrm, m2, n);
}

// constraint.py::_kbi
__device__ __noinline__ void kbi(const DModel* __restrict__ dm, const float* solref, const float* solimp, float pos,
                                    float* k, float* b, float* imp) {
  float timeconst = solref[0];
  const float dampratio = solref[1];
  const float dt2 = 2.f * dm->timestep;
  if (timeconst < dt2) timeconst = dt2;
  const float dmin = clipf(solimp[0], MJ_MINIMP, MJ_MAXIMP), dmax = clipf(solimp[1], MJ_MINIMP, MJ_MAXIMP);
  const float width = solimp[2] > MJ_MINVAL ? solimp[2] : MJ_MINVAL;
  const float mid = clipf(solimp[3], MJ_MINIMP, MJ_MAXIMP);
  const float power = solimp[4] > 1.f ? solimp[4] : 1.f;
  *k = 1.f / (dmax * dmax * timeconst * timeconst * dampratio * dampratio);
  *b = 2.f / (dmax * timeconst);
  if (solref[0] <= 0.f) *k = -solref[0] / (dmax * dmax);
  if (solref[1] <= 0.f) *b = -solref[1] / dmax;
  const float x = fabsf(pos) / width;
  if (x > 1.f) { *imp = dmax; return; }
  float y;
  if (x < mid) y = (1.f / pw(mid, power - 1.f)) * pw(x, power);
  else y = 1.f - (1.f / pw(1.f - mid, power - 1.f)) * pw(1.f - x, power);
  *imp = clipf(dmin + y * (dmax - dmin), dmin, dmax);
}

// collision_driver.py::collision in three warp passes:
//  0. one lane per geom: world pose into shared memory (region C);
//  1. one lane per geom pair: conservative bounding test (box centre + bounding radius against the plane / against the
//     other box's faces).  A culled pair is one the narrow phase would report no contact for (a separating face axis
//     exists), so the contact list is unchanged; survivors are compacted in pair order;
//  2. one half warp per surviving pair (two pairs side by side): cooperative narrow phase in registers, the active
//     contacts appended to the shared-memory contact list in (pair, slot) order.  Contacts that MJX would keep as zeroed rows
//     (dist >= 0) are dropped.  Returns ncon.
__device__ __noinline__ int collision(const DModel* __restrict__ dm, float* sm, int lane, int* status) {
  for (int g = lane; g < dm->ngeom; g += 32) {
    float* gp = sm + ar::GPOSE + g * 12;
    geom_pose(dm, sm, g, gp, gp + 3);
  }
  RSRX_SYNC();
  int* plist = reinterpret_cast<int*>(sm + ar::PLIST);
  int nsurv = 0;
  for (int base = 0; base < dm->npair; base += 32) {
    const int p = base + lane;
    bool keep = false;
    if (p < dm->npair) {
      const int g1 = dm->pair_g1[p], g2 = dm->pair_g2[p];
      const float slack = dm->pair_margin[p] + 1e-4f;
      const float* a = sm + ar::GPOSE + g1 * 12;
      const float* b = sm + ar::GPOSE + g2 * 12;
      const float t[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
      if (dm->geom_type[g1] == RSRX_GEOM_PLANE) {
        keep = a[3 + 2] * t[0] + a[3 + 5] * t[1] + a[3 + 8] * t[2] - dm->geom_rbound[g2] <= slack;
      } else {
        keep = true;
        const float r1 = dm->geom_rbound[g1] + slack, r2 = dm->geom_rbound[g2] + slack;
#pragma unroll
        for (int i = 0; i < 3; i++) {
          const float ta = a[3 + i] * t[0] + a[6 + i] * t[1] + a[9 + i] * t[2];
          const float tb = b[3 + i] * t[0] + b[6 + i] * t[1] + b[9 + i] * t[2];
          if (fabsf(ta) > dm->geom_size[g1][i] + r2 || fabsf(tb) > dm->geom_size[g2][i] + r1) keep = false;
        }
      }
    }
    const unsigned kept = __ballot_sync(0xffffffffu, keep);
    if (keep) plist[nsurv + __popc(kept & ((1u << lane) - 1u))] = p;
    nsurv += __popc(kept);
  }
  RSRX_SYNC();
#ifdef RSRX_PRINT_NSURV
  if (lane == 0) { printf("NSURV %d :", nsurv); for (int q = 0; q < nsurv; q++) printf(" %d-%d", dm->pair_g1[plist[q]], dm->pair_g2[plist[q]]); printf("\n"); }
#endif
  int ncon = 0, boff = 0;  // contacts so far / cursor into the Jacobian-row pool
  bool any_spill = false;
  Half hw;
  hw.shift = lane & 16; hw.l = lane & 15; hw.mask = 0xffffu << hw.shift;
#pragma unroll 1
  for (int sv = 0; sv < nsurv; sv += 2) {
    const int mine = sv + (lane >> 4);  // the pair this half works on
    const bool valid = mine < nsurv;
    const int p = plist[valid ? mine : sv];
    const int g1 = dm->pair_g1[p], g2 = dm->pair_g2[p];
    const float margin = dm->pair_margin[p];
    Manifold mf;
    mf.dist = 1.f;
    if (valid) {
      const float s2[3] = {dm->geom_size[g2][0], dm->geom_size[g2][1], dm->geom_size[g2][2]};
      const float* a = sm + ar::GPOSE + g1 * 12;
      const float* b = sm + ar::GPOSE + g2 * 12;
      if (dm->geom_type[g1] == RSRX_GEOM_PLANE) {
        plane_box(a, a + 3, b, b + 3, s2, hw, &mf);
      } else {
        const float s1[3] = {dm->geom_size[g1][0], dm->geom_size[g1][1], dm->geom_size[g1][2]};
        box_box(a, a + 3, s1, b, b + 3, s2, sm + ar::CSCRATCH + (lane >> 4) * 24, hw, &mf);
      }
    }
    __syncwarp();
    const bool active = valid && hw.l < 4 && (mf.dist - margin < 0.f);
    const unsigned am = __ballot_sync(0xffffffffu, active);
    if (am == 0u) continue;
    const int slot = ncon + __popc(am & ((1u << lane) - 1u));
    if (ncon + __popc(am) > dm->contact_cap) *status |= RSRX_STATUS_CONTACT_OVERFLOW;  // the caller hands the env-step to the large-capacity kernel
    // dof ranges of the two trees this pair joins (a static body contributes no columns) = width of its Jacobian rows
    const int b1 = dm->geom_bodyid[g1], b2 = dm->geom_bodyid[g2];
    const int t1 = dm->body_treeid[b1], t2 = dm->body_treeid[b2];
    const int na = dm->body_dofmask[b1] ? dm->tree_dofnum[t1] : 0, nb = dm->body_dofmask[b2] ? dm->tree_dofnum[t2] : 0;
    // row-pool allocation in contact order: half 0's contacts, then half 1's (a contact that does not fit is spilled
    // but still advances the cursor, so everything after it spills too: rare)
    const int sz = 4 * (na + nb);
    const int sz0 = __shfl_sync(0xffffffffu, sz, 0), sz1 = __shfl_sync(0xffffffffu, sz, 16);
    const int cnt0 = __popc(am & 0xffffu), cnt1 = __popc(am >> 16);
    const int rank = __popc(am & hw.mask & ((1u << lane) - 1u));
    const int off = boff + (hw.shift ? cnt0 * sz0 : 0) + rank * sz;
    if (active && slot < MAXC) {
      float frame[9];
      make_frame(frame, mf.nrm);
      float mu[3];
      const float* gfric_e = reinterpret_cast<const float* const*>(sm + ar::PTRS)[0];
      for (int i = 0; i < 3; i++)
        mu[i] = gfric_e ? fmaxf(gfric_e[g1 * 3 + i], gfric_e[g2 * 3 + i]) : fmaxf(dm->geom_friction[g1][i], dm->geom_friction[g2][i]);
      const int cols = (na ? dm->tree_dofadr[t1] : 0) | (na << 8) | ((nb ? dm->tree_dofadr[t2] : 0) << 16) | (nb << 24);
      float solref[2] = {dm->pair_solref[p][0], dm->pair_solref[p][1]}, solimp[5];
      for (int i = 0; i < 5; i++) solimp[i] = dm->pair_solimp[p][i];
      const float tran = dm->pair_tran[p];
      const float invw = (tran + mu[0] * mu[0] * tran) * 2.f * mu[0] * mu[0] / dm->impratio;
      float* cr = sm + ar::CON + slot * ar::CSTRIDE;
      float* ctm = sm + ar::CTMP + slot * ar::CTSTRIDE;
      for (int i = 0; i < 3; i++) ctm[ct::POS + i] = mf.pos[i];
      for (int i = 0; i < 9; i++) ctm[ct::FRAME + i] = frame[i];
      const float ps = mf.dist - margin;
      float k, bb, imp;
      kbi(dm, solref, solimp, ps, &k, &bb, &imp);
      float rr = invw * (1.f - imp) / imp;
      if (rr < MJ_MINVAL) rr = MJ_MINVAL;
      cr[cf::DIST] = mf.dist;
      cr[cf::MU] = mu[0]; cr[cf::MU + 1] = mu[0]; cr[cf::MU + 2] = mu[1];
      cr[cf::KIMPD] = k * imp * ps;
      cr[cf::B] = bb;
      cr[cf::D] = 1.f / rr;
      cr[cf::BODIES] = __int_as_float(b1 | (b2 << 8) | (g1 << 16) | (g2 << 24));
      cr[cf::COLS] = __int_as_float(cols);
      const bool fits = off + sz <= dm->pool_floats;
      any_spill |= !fits;
      cr[cf::BOFF] = __int_as_float(fits ? off : -(slot * 4 * NCOL + 1));
    }
    boff += cnt0 * sz0 + cnt1 * sz1;
    ncon += __popc(am);
  }
  any_spill = __any_sync(0xffffffffu, any_spill);
  if (lane == 0) reinterpret_cast<int*>(sm + ar::PTRS)[ar::FLAGS + 1] = any_spill ? 1 : 0;
  RSRX_SYNC();
  if (ncon > MAXC) ncon = MAXC;
  return ncon;
}

// the 4 x (na + nb) Jacobian base rows of a contact: in the shared-memory pool, or (overflow) in the env's spill row.
// SPILL = false is the instantiation for an env whose contacts all fit the pool (flag set by collision()): the pointer
// then provably stays in shared memory and the loads compile to LDS instead of generic LD (7 % of the step time).
template <bool SPILL>
__device__ __forceinline__ float* brow(float* sm, const float* cr) {
  const int off = __float_as_int(cr[cf::BOFF]);
  if (!SPILL) return sm + ar::BROW + off;
  return off >= 0 ? sm + ar::BROW + off : reinterpret_cast<float* const*>(sm + ar::PTRS)[3] + (-off - 1);
}
__device__ __forceinline__ bool rows_spilled(const float* sm) {
  return reinterpret_cast<const int*>(sm + ar::PTRS)[ar::FLAGS + 1] != 0;
}

// UB[c][p] = B[c][p][:] . x over the contact's dof columns (x: nv-vector in shared memory)
template <bool SPILL>
__device__ __forceinline__ void mul_B_rows(float* sm, int lane, int ncon, const float* x) {
#pragma unroll 1
  for (int t = lane; t < ncon * 4; t += 32) {
    const int c = t >> 2;
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const int cols = __float_as_int(cr[cf::COLS]);
    const int a0 = cols & 0xff, na = (cols >> 8) & 0xff, b0 = (cols >> 16) & 0xff, nb = (cols >> 24) & 0xff;
    const float* Bp = brow<SPILL>(sm, cr) + (t & 3) * (na + nb);
    float s = 0.f;
#pragma unroll 1
    for (int i = 0; i < na; i++) s += Bp[i] * x[a0 + i];
#pragma unroll 1
    for (int i = 0; i < nb; i++) s += Bp[na + i] * x[b0 + i];
    sm[ar::UB + t] = s;
  }
}
__device__ __noinline__ void mul_B(float* sm, int lane, int ncon, const float* x) {
  if (rows_spilled(sm)) mul_B_rows<true>(sm, lane, ncon, x);
  else mul_B_rows<false>(sm, lane, ncon, x);
  RSRX_SYNC();
}

// D of constraint row r
__device__ __forceinline__ float row_D(const float* sm, int r, int nsr) {
  return r < nsr ? sm[ar::E_DS + r] : sm[ar::CON + ((r - nsr) / 6) * ar::CSTRIDE + cf::D];
}

// ----------------------------------------------------------------- constraints
// constraint.py::make_constraint.  Sparse rows (equality, dof friction, joint
// limits) are kept as (dof, coef) pairs; contact rows as 4 base rows per contact
// (normal, tangent1, tangent2, torsion) of the contact-frame Jacobian difference,
// from which the 6 pyramid-edge rows J_n +- mu_k J_k are formed on the fly.
// Row order: [sparse rows][6 rows per contact].  Returns nsr (number of sparse rows).
__device__ __noinline__ int make_constraint(const DModel* __restrict__ dm, float* sm, int lane, int ncon) {
  const int nv = dm->nv;
  int* sr_dofa = reinterpret_cast<int*>(sm + ar::SR_DOFA);
  int* sr_dofb = reinterpret_cast<int*>(sm + ar::SR_DOFB);
  int* sr_type = reinterpret_cast<int*>(sm + ar::SR_TYPE);
  // --- sparse rows: lane L handles candidate L: [eq | friction dofs | limited joints]
  const int neq = dm->neq, nfr_c = nv, nlim_c = dm->njnt;
  const int cand = lane;
  bool have = false;
  int type = 0, dofa = 0, dofb = -1;
  float ca = 0.f, cb = 0.f, pos = 0.f, invw = 0.f, floss = 0.f, margin = 0.f;
  float solref[2] = {0.02f, 1.f}, solimp[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.f};
  if (cand < neq) {
    const int e = cand, q1 = dm->eq_q1[e], q2 = dm->eq_q2[e];
    have = true; type = 0; dofa = dm->eq_d1[e]; ca = 1.f;
    const float pos1 = sm[ar::QPOS + q1] - dm->qpos0[q1];
    if (q2 >= 0) {
      const float dif = sm[ar::QPOS + q2] - dm->qpos0[q2];
      const float dp[5] = {1.f, dif, dif * dif, dif * dif * dif, dif * dif * dif * dif};
      float deriv = 0.f, poly = 0.f;
      for (int i = 0; i < 5; i++) poly += dm->eq_data[e][i] * dp[i];
      for (int i = 1; i < 5; i++) deriv += dm->eq_data[e][i] * dp[i - 1] * (float)i;
      dofb = dm->eq_d2[e]; cb = -deriv;
      pos = pos1 - poly;
    } else {
      pos = pos1 - dm->eq_data[e][0];
    }
    invw = dm->eq_invweight[e];
    solref[0] = dm->eq_solref[e][0]; solref[1] = dm->eq_solref[e][1];
    for (int i = 0; i < 5; i++) solimp[i] = dm->eq_solimp[e][i];
  } else if (cand < neq + nfr_c) {
    const int d = cand - neq;
    if (dm->dof_hasfriction[d]) {
      have = true; type = 1; dofa = d; ca = 1.f; pos = 0.f; invw = dm->dof_invweight0[d];
      const float* floss_e = reinterpret_cast<const float* const*>(sm + ar::PTRS)[2];
      floss = floss_e ? floss_e[d] : dm->dof_frictionloss[d];
      solref[0] = dm->dof_solref[d][0]; solref[1] = dm->dof_solref[d][1];
      for (int i = 0; i < 5; i++) solimp[i] = dm->dof_solimp[d][i];
    }
  }
  // limits may not fit the first 32 candidates: handled by a second candidate slot
  bool have2 = false;
  int dofa2 = 0;
  float ca2 = 0.f, pos2 = 0.f, invw2 = 0.f, margin2 = 0.f;
  float solref2[2] = {0.02f, 1.f}, solimp2[5] = {0.9f, 0.95f, 0.001f, 0.5f, 2.f};
  if (lane < nlim_c) {
    const int j = lane;
    if (dm->jnt_limited[j] && dm->jnt_type[j] != RSRX_JNT_FREE) {
      const float q = sm[ar::QPOS + dm->jnt_qposadr[j]];
      const float dmin = q - dm->jnt_range[j][0], dmax = dm->jnt_range[j][1] - q;
      margin2 = dm->jnt_margin[j];
      const float ps = fminf(dmin, dmax) - margin2;
      if (ps < 0.f) {
        have2 = true; dofa2 = dm->jnt_dofadr[j]; ca2 = dmin < dmax ? 1.f : -1.f; pos2 = ps;
        invw2 = dm->dof_invweight0[dofa2];
        solref2[0] = dm->jnt_solref[j][0]; solref2[1] = dm->jnt_solref[j][1];
        for (int i = 0; i < 5; i++) solimp2[i] = dm->jnt_solimp[j][i];
      }
    }
  }
  const unsigned m1 = __ballot_sync(0xffffffffu, have), m2 = __ballot_sync(0xffffffffu, have2);
  const int n1 = __popc(m1), nsr = n1 + __popc(m2);
  const unsigned lt = (1u << lane) - 1u;
  if (have) {
    const int r = __popc(m1 & lt);
    float k, b, imp;
    kbi(dm, solref, solimp, pos - margin, &k, &b, &imp);
    float rr = invw * (1.f - imp) / imp;
    if (rr < MJ_MINVAL) rr = MJ_MINVAL;
    float vel = ca * sm[ar::QVEL + dofa];
    if (dofb >= 0) vel += cb * sm[ar::QVEL + dofb];
    // oracle sums J[d]*qvel[d] in dof order
    if (dofb >= 0 && dofb < dofa) vel = cb * sm[ar::QVEL + dofb] + ca * sm[ar::QVEL + dofa];
    sm[ar::E_DS + r] = 1.f / rr;
    sm[ar::E_AREF + r] = -b * vel - k * imp * (pos - margin);
    sr_dofa[r] = dofa; sr_dofb[r] = dofb; sr_type[r] = type;
    sm[ar::SR_CA + r] = ca; sm[ar::SR_CB + r] = cb; sm[ar::SR_FLOSS + r] = floss; sm[ar::SR_RF + r] = floss * rr;
  }
  if (have2) {
    const int r = n1 + __popc(m2 & lt);
    float k, b, imp;
    kbi(dm, solref2, solimp2, pos2, &k, &b, &imp);
    float rr = invw2 * (1.f - imp) / imp;
    if (rr < MJ_MINVAL) rr = MJ_MINVAL;
    const float vel = ca2 * sm[ar::QVEL + dofa2];
    sm[ar::E_DS + r] = 1.f / rr;
    sm[ar::E_AREF + r] = -b * vel - k * imp * pos2;
    sr_dofa[r] = dofa2; sr_dofb[r] = -1; sr_type[r] = 2;
    sm[ar::SR_CA + r] = ca2; sm[ar::SR_CB + r] = 0.f; sm[ar::SR_FLOSS + r] = 0.f; sm[ar::SR_RF + r] = 0.f;
  }
  // --- contact base rows B[c][p][col] (support.jac + frame rotation); column -> dof through cf::COLS
#pragma unroll 1
  for (int t = lane; t < ncon * NCOL; t += 32) {
    const int c = t / NCOL, col = t - c * NCOL;
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const float* ctm = sm + ar::CTMP + c * ar::CTSTRIDE;
    const int cols = __float_as_int(cr[cf::COLS]);
    const int a0 = cols & 0xff, na = (cols >> 8) & 0xff, b0 = (cols >> 16) & 0xff, nb = (cols >> 24) & 0xff;
    const int w = na + nb;  // row width of this contact
    if (col >= w) continue;
    float* B = brow<true>(sm, cr);  // (one-shot: the generic pointer is fine here)
    const int d = col < na ? a0 + col : b0 + col - na;
    const int bodies = __float_as_int(cr[cf::BODIES]);
    const int b1 = bodies & 0xff, b2 = (bodies >> 8) & 0xff;
    const bool in1 = (dm->body_dofmask[b1] >> d) & 1u, in2 = (dm->body_dofmask[b2] >> d) & 1u;
    float dp[3] = {0.f, 0.f, 0.f}, dr[3] = {0.f, 0.f, 0.f};
    const float* cd = sm + ar::CDOF + d * 6;
    if (in2) {
      const float* sc = sm + ar::SCOM + dm->body_treeid[b2] * 3;
      float off[3] = {ctm[ct::POS] - sc[0], ctm[ct::POS + 1] - sc[1], ctm[ct::POS + 2] - sc[2]}, x[3];
      cross3(x, cd, off);
      for (int i = 0; i < 3; i++) { dp[i] = cd[3 + i] + x[i]; dr[i] = cd[i]; }
    }
    if (in1) {
      const float* sc = sm + ar::SCOM + dm->body_treeid[b1] * 3;
      float off[3] = {ctm[ct::POS] - sc[0], ctm[ct::POS + 1] - sc[1], ctm[ct::POS + 2] - sc[2]}, x[3];
      cross3(x, cd, off);
      for (int i = 0; i < 3; i++) { dp[i] -= cd[3 + i] + x[i]; dr[i] -= cd[i]; }
    }
    B[col] = dot3(ctm + ct::FRAME, dp);
    B[w + col] = dot3(ctm + ct::FRAME + 3, dp);
    B[2 * w + col] = dot3(ctm + ct::FRAME + 6, dp);
    B[3 * w + col] = dot3(ctm + ct::FRAME, dr);
  }
  RSRX_SYNC();
  {  // does any contact couple two kinematic trees?  (selects the Cholesky block structure of H)
    bool cpl = false;
    for (int c = lane; c < ncon; c += 32) {
      const int cols = __float_as_int(sm[ar::CON + c * ar::CSTRIDE + cf::COLS]);
      cpl |= ((cols >> 8) & 0xff) != 0 && ((cols >> 24) & 0xff) != 0;
    }
    const bool any = __any_sync(0xffffffffu, cpl);
    if (lane == 0) reinterpret_cast<int*>(sm + ar::PTRS)[ar::FLAGS] = any ? 1 : 0;
  }
  // --- contact rows: aref (D lives in the contact record)
  mul_B(sm, lane, ncon, sm + ar::QVEL);  // base velocities u_p = B_p . qvel -> UB
#pragma unroll 1
  for (int t = lane; t < ncon * 6; t += 32) {
    const int c = t / 6, e = t - c * 6, k = e >> 1;
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const float f = (e & 1) ? -cr[cf::MU + k] : cr[cf::MU + k];
    const float vel = sm[ar::UB + c * 4] + sm[ar::UB + c * 4 + 1 + k] * f;
    sm[ar::E_AREF + nsr + t] = -cr[cf::B] * vel - cr[cf::KIMPD];
  }
  RSRX_SYNC();
  return nsr;
}

// ------------------------------------------------------------ velocity / forces
// smooth.py::com_vel, passive.py, smooth.py::rne, forward.py::fwd_actuation,
// fwd_acceleration
__device__ __noinline__ void velocity_and_forces(const DModel* __restrict__ dm, float* sm, int lane) {
  const int nv = dm->nv, nbody = dm->nbody;
  float* cacc = sm + ar::CRB;           // [NB][6] (crb no longer needed)
  // com_vel: level by level
  if (lane < nbody && dm->body_static[lane]) {
    for (int i = 0; i < 6; i++) { sm[ar::CVEL + lane * 6 + i] = 0.f; cacc[lane * 6 + i] = (i >= 3) ? -dm->gravity[i - 3] : 0.f; }
  }
  RSRX_SYNC();
  for (int lev = 1; lev < dm->nlevel; ++lev) {
    const int b = lane;
    if (b < nbody && dm->body_depth[b] == lev && !dm->body_static[b]) {
      const int p = dm->body_parentid[b];
      float cvel[6];
      for (int i = 0; i < 6; i++) cvel[i] = sm[ar::CVEL + p * 6 + i];
      const int jn = dm->body_jntnum[b], ja0 = dm->body_jntadr[b];
      for (int k = 0; k < jn; k++) {
        const int j = ja0 + k, d = dm->jnt_dofadr[j];
        if (dm->jnt_type[j] == RSRX_JNT_FREE) {
          for (int a = 0; a < 3; a++)
            for (int i = 0; i < 6; i++) cvel[i] += sm[ar::CDOF + (d + a) * 6 + i] * sm[ar::QVEL + d + a];
          for (int a = 0; a < 3; a++) {
            for (int i = 0; i < 6; i++) sm[ar::CDOFDOT + (d + a) * 6 + i] = 0.f;
            float r[6];
            motion_cross(r, cvel, sm + ar::CDOF + (d + 3 + a) * 6);
            for (int i = 0; i < 6; i++) sm[ar::CDOFDOT + (d + 3 + a) * 6 + i] = r[i];
          }
          for (int a = 3; a < 6; a++)
            for (int i = 0; i < 6; i++) cvel[i] += sm[ar::CDOF + (d + a) * 6 + i] * sm[ar::QVEL + d + a];
        } else {
          float r[6];
          motion_cross(r, cvel, sm + ar::CDOF + d * 6);
          for (int i = 0; i < 6; i++) sm[ar::CDOFDOT + d * 6 + i] = r[i];
          for (int i = 0; i < 6; i++) cvel[i] += sm[ar::CDOF + d * 6 + i] * sm[ar::QVEL + d];
        }
      }
      for (int i = 0; i < 6; i++) sm[ar::CVEL + b * 6 + i] = cvel[i];
      // rne forward pass: cacc
      float ca[6];
      for (int i = 0; i < 6; i++) ca[i] = cacc[p * 6 + i];
      const int dn = dm->body_dofnum[b], d0 = dm->body_dofadr[b];
      for (int k = 0; k < dn; k++)
        for (int i = 0; i < 6; i++) ca[i] += sm[ar::CDOFDOT + (d0 + k) * 6 + i] * sm[ar::QVEL + d0 + k];
      for (int i = 0; i < 6; i++) cacc[b * 6 + i] = ca[i];
    }
    RSRX_SYNC();
  }
  // local body forces -> HH scratch [NB][6] (H is rebuilt later by the solver)
  float* lfrc = sm + ar::HH;
  if (lane < nbody && lane > 0 && !dm->body_static[lane]) {
    const int b = lane;
    float f1[6], f2[6], f3[6];
    inert_mul(f1, sm + ar::CINERT + b * 10, cacc + b * 6);
    inert_mul(f2, sm + ar::CINERT + b * 10, sm + ar::CVEL + b * 6);
    motion_cross_force(f3, sm + ar::CVEL + b * 6, f2);
    for (int i = 0; i < 6; i++) lfrc[b * 6 + i] = f1[i] + f3[i];
  }
  RSRX_SYNC();
  // qfrc_bias[d] = cdof[d] . sum_{c in subtree(body(d))} lfrc[c]; passive; actuation
  if (lane < nv) {
    const int d = lane, b = dm->dof_bodyid[d];
    float acc[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    for (int c = dm->body_subtree_end[b] - 1; c >= b; --c)
      for (int i = 0; i < 6; i++) acc[i] += lfrc[c * 6 + i];
    float s = 0.f;
    for (int i = 0; i < 6; i++) s += sm[ar::CDOF + d * 6 + i] * acc[i];
    sm[ar::V_BIAS + d] = s;
    sm[ar::V_ACT + d] = 0.f;
  }
  RSRX_SYNC();
  if (lane < dm->nu) {
    const int u = lane, qa = dm->act_qadr[u], d = dm->act_dof[u];
    float ctrl = sm[ar::CTRL + u];
    if (dm->act_ctrllimited[u]) ctrl = clipf(ctrl, dm->act_ctrlrange[u][0], dm->act_ctrlrange[u][1]);
    const float gear = dm->act_gear[u];
    const float length = gear * sm[ar::QPOS + qa], velocity = gear * sm[ar::QVEL + d];
    const float bias = dm->act_bias[u][0] + dm->act_bias[u][1] * length + dm->act_bias[u][2] * velocity;
    float force = dm->act_gain[u] * ctrl + bias;
    if (dm->act_forcelimited[u]) force = clipf(force, dm->act_forcerange[u][0], dm->act_forcerange[u][1]);
    atomicAdd(sm + ar::V_ACT + d, gear * force);
  }
  RSRX_SYNC();
  if (lane < nv) {
    const int d = lane;
    float fa = sm[ar::V_ACT + d];
    if (dm->dof_actfrclimited[d]) fa = clipf(fa, dm->dof_actfrcrange[d][0], dm->dof_actfrcrange[d][1]);
    sm[ar::V_ACT + d] = fa;
    const float fs = -sm[ar::DAMP + d] * sm[ar::QVEL + d] - sm[ar::V_BIAS + d] + fa;
    sm[ar::V_SMOOTH + d] = fs;
    sm[ar::V_QACCS + d] = fs;
  }
  // factor_m + solve_m: qacc_smooth = M^-1 qfrc_smooth (factor a scratch copy of M; H is rebuilt later)
  copy_M_permuted(dm, sm, lane, nullptr, 0.f);
  warp_chol_factor(dm, sm, lane, true);
  warp_chol_solve(dm, sm, sm + ar::V_QACCS, lane, true);
}

// ---------------------------------------------------------------------- solver
struct SolverDims { int nsr, ncon, nrow; };

// out[r] = J[r] . x for every row (x: nv-vector in shared memory)
__device__ __noinline__ void mul_J(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon, const float* x,
                                   float* out) {
  const int* sr_dofa = reinterpret_cast<const int*>(sm + ar::SR_DOFA);
  const int* sr_dofb = reinterpret_cast<const int*>(sm + ar::SR_DOFB);
  if (lane < nsr) {
    const int a = sr_dofa[lane], b = sr_dofb[lane];
    float s = sm[ar::SR_CA + lane] * x[a];
    if (b >= 0) s = (b < a) ? sm[ar::SR_CB + lane] * x[b] + s : s + sm[ar::SR_CB + lane] * x[b];
    out[lane] = s;
  }
  mul_B(sm, lane, ncon, x);
#pragma unroll 1
  for (int t = lane; t < ncon * 6; t += 32) {
    const int c = t / 6, e = t - c * 6, k = e >> 1;
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const float f = (e & 1) ? -cr[cf::MU + k] : cr[cf::MU + k];
    out[nsr + t] = sm[ar::UB + c * 4] + sm[ar::UB + c * 4 + 1 + k] * f;
  }
  RSRX_SYNC();
}

__device__ __noinline__ void mul_M(const DModel* __restrict__ dm, const float* sm, int lane, const float* x, float* out) {
  const int nv = dm->nv;
  if (lane < nv) {
    float s = 0.f;
    // M couples only dofs of the same kinematic tree: the other entries are exact zeros and are skipped
    const int j1 = dm->dof_tree_hi[lane];
#pragma unroll 4
    for (int j = dm->dof_tree_lo[lane]; j <= j1; j++) {
      const int hi = lane > j ? lane : j, lo = lane > j ? j : lane;
      s += sm[ar::MM + ((hi * (hi + 1)) >> 1) + lo] * x[j];
    }
    out[lane] = s;
  }
}

// Per-row piecewise-quadratic description shared by _update_constraint and the
// line search: a row is in its quadratic zone iff lo < x < hi (x = Jaref along the
// search direction); outside, a dof-friction row is linear with force -+floss,
// every other row contributes nothing.  equality: (-inf, inf); friction:
// (-R*floss, R*floss); limit / contact: (-inf, 0).
struct RowShape { float lo, hi, fl, rf; };
__device__ __forceinline__ RowShape row_shape(const float* sm, int r, int nsr) {
  RowShape s;
  s.lo = -INFINITY; s.hi = 0.f; s.fl = 0.f; s.rf = 0.f;
  if (r < nsr) {
    const int type = reinterpret_cast<const int*>(sm + ar::SR_TYPE)[r];
    if (type == 0) s.hi = INFINITY;
    else if (type == 1) { s.fl = sm[ar::SR_FLOSS + r]; s.rf = sm[ar::SR_RF + r]; s.lo = -s.rf; s.hi = s.rf; }
  }
  return s;
}

// sum over contacts of B_c[:, d] . g_c (dof d's share of the contact forces held in UB)
template <bool SPILL>
__device__ __forceinline__ float contact_qfrc(float* sm, int d, int ncon) {
  float s = 0.f;
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const int cols = __float_as_int(cr[cf::COLS]);
    const int a0 = cols & 0xff, na = (cols >> 8) & 0xff, b0 = (cols >> 16) & 0xff, nb = (cols >> 24) & 0xff;
    const int col = (d >= a0 && d < a0 + na) ? d - a0 : ((d >= b0 && d < b0 + nb) ? na + d - b0 : -1);
    if (col < 0) continue;
    const float* B = brow<SPILL>(sm, cr) + col;
    const float* g = sm + ar::UB + c * 4;
    const int w = na + nb;
    s += B[0] * g[0] + B[w] * g[1] + B[2 * w] * g[2] + B[3 * w] * g[3];
  }
  return s;
}

// solver.py::_update_constraint.  Returns the total cost; writes E_ACT, qfrc_constraint.
__device__ __noinline__ float update_constraint(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon,
                                                float* gauss_out, bool* changed_out) {
  const int nv = dm->nv, nrow = nsr + 6 * ncon;
  const bool spilled = rows_spilled(sm);
  float cost = 0.f;
  bool changed = false;
  unsigned* actw = reinterpret_cast<unsigned*>(sm + ar::E_ACT);  // active flags, one bit per row
#pragma unroll 1
  for (int base = 0; base < nrow; base += 32) {
    const int r = base + lane;
    bool quad = false;
    if (r < nrow) {
      const float ja = sm[ar::E_JAREF + r], D = row_D(sm, r, nsr);
      const RowShape s = row_shape(sm, r, nsr);
      quad = ja > s.lo && ja < s.hi;
      const bool below = ja <= s.lo;
      const float f = quad ? -D * ja : (below ? s.fl : -s.fl);
      cost += quad ? 0.5f * D * ja * ja : s.fl * (-0.5f * s.rf + (below ? -ja : ja));
      sm[ar::E_JV + r] = f;  // E_JV doubles as the force array between line searches
    }
    const unsigned word = __ballot_sync(0xffffffffu, quad);
    changed |= actw[base >> 5] != word;
    RSRX_SYNC();
    if (lane == 0) actw[base >> 5] = word;
  }
  *changed_out = changed;
  RSRX_SYNC();
  // contact forces in base-row space: g0 = sum f, g_{1+k} = mu_k (f_{2k} - f_{2k+1})
#pragma unroll 1
  for (int t = lane; t < ncon * 4; t += 32) {
    const int c = t >> 2, p = t & 3;
    const float* fr = sm + ar::E_JV + nsr + c * 6;
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    float g;
    if (p == 0) g = fr[0] + fr[1] + fr[2] + fr[3] + fr[4] + fr[5];
    else g = cr[cf::MU + p - 1] * (fr[2 * (p - 1)] - fr[2 * (p - 1) + 1]);
    sm[ar::UB + t] = g;
  }
  if (lane < nv) sm[ar::V_QFRCC + lane] = 0.f;
  RSRX_SYNC();
  if (lane < nsr) {  // sparse rows: at most two dofs each
    const int a = reinterpret_cast<const int*>(sm + ar::SR_DOFA)[lane], b = reinterpret_cast<const int*>(sm + ar::SR_DOFB)[lane];
    const float f = sm[ar::E_JV + lane];
    atomicAdd(sm + ar::V_QFRCC + a, sm[ar::SR_CA + lane] * f);
    if (b >= 0) atomicAdd(sm + ar::V_QFRCC + b, sm[ar::SR_CB + lane] * f);
  }
  RSRX_SYNC();
  float gpart = 0.f;
  if (lane < nv) {
    const int d = lane;
    const float s = sm[ar::V_QFRCC + d] + (spilled ? contact_qfrc<true>(sm, d, ncon) : contact_qfrc<false>(sm, d, ncon));
    sm[ar::V_QFRCC + d] = s;
    gpart = (sm[ar::V_MA + d] - sm[ar::V_SMOOTH + d]) * (sm[ar::V_QACC + d] - sm[ar::V_QACCS + d]);
  }
  const float gauss = 0.5f * warp_sum(gpart);
  cost = warp_sum(cost) + gauss;
  *gauss_out = gauss;
  RSRX_SYNC();
  return cost;
}

// H += sum_c J_c^T diag(D active) J_c, contact by contact: the lanes take the w (w + 1) / 2 column pairs of the
// contact's w = na + nb Jacobian columns (packed lower-triangle order, DModel::tri_i / tri_j) and add the 4 x 4
// base-row Gram form of the pair to its H entry.  Every pair is a structural non-zero; no (entry, contact) misses.
template <bool SPILL>
__device__ __forceinline__ void contact_hessian_scatter(const DModel* __restrict__ dm, float* sm, int lane, int ncon) {
#pragma unroll 1
  for (int c = 0; c < ncon; c++) {
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const int cols = __float_as_int(cr[cf::COLS]);
    const int a0 = cols & 0xff, na = (cols >> 8) & 0xff, b0 = (cols >> 16) & 0xff, nb = (cols >> 24) & 0xff;
    const int w = na + nb, npair = (w * (w + 1)) >> 1;
    const float* B = brow<SPILL>(sm, cr);
    const float* W = sm + ar::CW + c * 8;
#pragma unroll 1
    for (int idx = lane; idx < npair; idx += 32) {
      const int ci = dm->tri_i[idx], cj = dm->tri_j[idx];
      const float b0i = B[ci], b0j = B[cj];
      float acc = W[0] * b0i * b0j;
#pragma unroll
      for (int k = 0; k < 3; k++) {
        const float bki = B[(1 + k) * w + ci], bkj = B[(1 + k) * w + cj];
        acc += W[1 + k] * (b0i * bkj + bki * b0j) + W[4 + k] * bki * bkj;
      }
      const int pi = dm->pos_of_dof[ci < na ? a0 + ci : b0 + ci - na], pj = dm->pos_of_dof[cj < na ? a0 + cj : b0 + cj - na];
      atomicAdd(sm + ar::HH + tri(max(pi, pj)) + min(pi, pj), acc);
    }
  }
}

// solver.py::_update_gradient (Newton): grad, H = M + J^T diag(D*active) J (in
// the block-permuted dof order), Cholesky, Mgrad = H^-1 grad
__device__ __noinline__ void update_gradient(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon,
                                             bool reuse_factor) {
  const bool tree_blocks = reinterpret_cast<const int*>(sm + ar::PTRS)[ar::FLAGS] == 0;  // set by make_constraint
  const bool spilled = rows_spilled(sm);
  // (grad itself is formed by the caller, which tests convergence on it before asking for the Newton direction)
  if (reuse_factor) {  // same active set as the previous iteration: H, hence its factor in ar::HH, is unchanged
    RSRX_SYNC();
    warp_chol_solve(dm, sm, sm + ar::V_MGRAD, lane, tree_blocks);
    return;
  }
  copy_M_permuted(dm, sm, lane, nullptr, 0.f);
  // per-contact weights of the 4x4 base-row Gram form: W00 = sum_r w_r,
  // U_k = (w_2k - w_2k+1) mu_k, V_k = (w_2k + w_2k+1) mu_k^2, w_r = D * active_r
#pragma unroll 1
  for (int c = lane; c < ncon; c += 32) {
    const float* cr = sm + ar::CON + c * ar::CSTRIDE;
    const unsigned* actw = reinterpret_cast<const unsigned*>(sm + ar::E_ACT);
    const int r0 = nsr + c * 6;
    const float D = cr[cf::D];
    float w00 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      const int rp = r0 + 2 * k, rm = rp + 1;
      const float wp = ((actw[rp >> 5] >> (rp & 31)) & 1u) ? D : 0.f, wm = ((actw[rm >> 5] >> (rm & 31)) & 1u) ? D : 0.f;
      const float mu = cr[cf::MU + k];
      w00 += wp + wm;
      sm[ar::CW + c * 8 + 1 + k] = (wp - wm) * mu;
      sm[ar::CW + c * 8 + 4 + k] = (wp + wm) * mu * mu;
    }
    sm[ar::CW + c * 8] = w00;
  }
  RSRX_SYNC();
  if (lane < nsr && ((reinterpret_cast<const unsigned*>(sm + ar::E_ACT)[0] >> lane) & 1u)) {  // sparse rows touch one diagonal entry (equality: a 2x2 block)
    const int a = reinterpret_cast<const int*>(sm + ar::SR_DOFA)[lane], b = reinterpret_cast<const int*>(sm + ar::SR_DOFB)[lane];
    const float ca = sm[ar::SR_CA + lane], cb = sm[ar::SR_CB + lane], D = sm[ar::E_DS + lane];
    const int pa = dm->pos_of_dof[a];
    atomicAdd(sm + ar::HH + tri(pa) + pa, ca * D * ca);
    if (b >= 0) {
      const int pb = dm->pos_of_dof[b];
      atomicAdd(sm + ar::HH + tri(pb) + pb, cb * D * cb);
      atomicAdd(sm + ar::HH + tri(max(pa, pb)) + min(pa, pb), ca * D * cb);
    }
  }
  RSRX_SYNC();
  if (spilled) contact_hessian_scatter<true>(dm, sm, lane, ncon);
  else contact_hessian_scatter<false>(dm, sm, lane, ncon);
  RSRX_SYNC();
  warp_chol_factor(dm, sm, lane, tree_blocks);
  warp_chol_solve(dm, sm, sm + ar::V_MGRAD, lane, tree_blocks);
}

// _Context.create: qacc <- src, Jaref, Ma, constraint update.  Returns cost.
__device__ __noinline__ float ctx_create(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon,
                                         const float* src, float* gauss) {
  const int nv = dm->nv, nrow = nsr + 6 * ncon;
  if (lane < nv) sm[ar::V_QACC + lane] = src[lane];
  RSRX_SYNC();
  mul_J(dm, sm, lane, nsr, ncon, sm + ar::V_QACC, sm + ar::E_JAREF);
#pragma unroll 1
  for (int r = lane; r < nrow; r += 32) sm[ar::E_JAREF + r] -= sm[ar::E_AREF + r];
  mul_M(dm, sm, lane, sm + ar::V_QACC, sm + ar::V_MA);
  RSRX_SYNC();
  bool changed;
  return update_constraint(dm, sm, lane, nsr, ncon, gauss, &changed);
}

struct LSPoint { float alpha, cost, d0, d1; };

// solver.py::_linesearch with _LSPoint.create evaluated for three alphas per pass
// over the rows.  The loop body is kept small enough for the L0 instruction cache
// (the kernel is instruction-fetch bound): sparse rows (one per lane) live in
// registers in the general piecewise form, contact rows (one-sided quadratics) are
// streamed from shared memory, the nine partial sums are reduced by one rolled
// butterfly, and the two start-up points go through the same evaluation code.
// Returns the number of bracketing iterations.
__device__ __noinline__ int linesearch(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon, float gauss) {
  const int nv = dm->nv, nrow = nsr + 6 * ncon;
  float s2 = 0.f, g1 = 0.f, g2 = 0.f;
  mul_M(dm, sm, lane, sm + ar::V_SEARCH, sm + ar::V_MV);
  RSRX_SYNC();
  if (lane < nv) {
    const float s = sm[ar::V_SEARCH + lane];
    s2 = s * s;
    g1 = s * sm[ar::V_MA + lane] - s * sm[ar::V_SMOOTH + lane];
    g2 = s * sm[ar::V_MV + lane];
  }
  const float smag = sqrtf(warp_sum(s2)) * dm->meaninertia * (float)(nv > 1 ? nv : 1);
  const float gtol = dm->tolerance * dm->ls_tolerance * smag;
  const float qg0 = gauss, qg1 = warp_sum(g1), qg2 = 0.5f * warp_sum(g2);
  mul_J(dm, sm, lane, nsr, ncon, sm + ar::V_SEARCH, sm + ar::E_JV);
  // sparse row of this lane, general piecewise form (zero contribution when lane >= nsr)
  float s_ja = 0.f, s_jv = 0.f, s_c0 = 0.f, s_c1 = 0.f, s_c2 = 0.f, s_lm = 0.f, s_lp = 0.f, s_l1 = 0.f;
  float s_lo = 0.f, s_hi = 0.f;  // empty quadratic zone
  if (lane < nsr) {
    const float ja = sm[ar::E_JAREF + lane], jv = sm[ar::E_JV + lane], D = sm[ar::E_DS + lane];
    const RowShape s = row_shape(sm, lane, nsr);
    s_ja = ja; s_jv = jv; s_lo = s.lo; s_hi = s.hi;
    s_c0 = 0.5f * ja * ja * D; s_c1 = jv * ja * D; s_c2 = 0.5f * jv * jv * D;
    s_lm = s.fl * (-0.5f * s.rf - ja); s_lp = s.fl * (-0.5f * s.rf + ja); s_l1 = s.fl * jv;
  }
  LSPoint p0, lo, hi;
  p0.alpha = p0.cost = p0.d0 = p0.d1 = 0.f;
  lo = hi = p0;
  float anchor_lo = -1e30f, anchor_hi = -1e30f;  // cycle detection, see below
  int anchor_it = 0;
  bool jumped = false;
  float al[3] = {0.f, 0.f, 0.f};
  int phase = 0, it = 0;
#pragma unroll 1
  for (;;) {
    // ---- _LSPoint.create x 3
    float q[9];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      const float x = __fadd_rn(s_ja, __fmul_rn(al[a], s_jv));
      const bool quad = x > s_lo && x < s_hi, below = x <= s_lo;
      q[3 * a] = quad ? s_c0 : (below ? s_lm : s_lp);
      q[3 * a + 1] = quad ? s_c1 : (below ? -s_l1 : s_l1);
      q[3 * a + 2] = quad ? s_c2 : 0.f;
    }
#pragma unroll 1
    for (int r = nsr + lane; r < nrow; r += 32) {
      const float ja = sm[ar::E_JAREF + r], jv = sm[ar::E_JV + r], D = row_D(sm, r, nsr);
      const float c0 = 0.5f * ja * ja * D, c1 = jv * ja * D, c2 = 0.5f * jv * jv * D;
#pragma unroll
      for (int a = 0; a < 3; a++) {
        const float x = __fadd_rn(ja, __fmul_rn(al[a], jv));
        if (x < 0.f) { q[3 * a] += c0; q[3 * a + 1] += c1; q[3 * a + 2] += c2; }
      }
    }
    // warp sums of the nine partials in 22 shuffles instead of 45: q[0..7] by recursive halving (each step a lane
    // keeps half of its values and hands the other half to its partner), broadcast back from lanes 0, 4, .., 28;
    // q[8] by a plain butterfly.  Every lane ends up with the same bits.
    {
      constexpr unsigned FULL = 0xffffffffu;
      const bool b4 = lane & 16, b3 = lane & 8, b2 = lane & 4;
      float r4[4], r2[2];
#pragma unroll
      for (int i = 0; i < 4; i++) r4[i] = (b4 ? q[i + 4] : q[i]) + __shfl_xor_sync(FULL, b4 ? q[i] : q[i + 4], 16);
#pragma unroll
      for (int i = 0; i < 2; i++) r2[i] = (b3 ? r4[i + 2] : r4[i]) + __shfl_xor_sync(FULL, b3 ? r4[i] : r4[i + 2], 8);
      float r1 = (b2 ? r2[1] : r2[0]) + __shfl_xor_sync(FULL, b2 ? r2[0] : r2[1], 4);
      r1 += __shfl_xor_sync(FULL, r1, 2);
      r1 += __shfl_xor_sync(FULL, r1, 1);
      float q8 = q[8];
#pragma unroll
      for (int o = 16; o; o >>= 1) q8 += __shfl_xor_sync(FULL, q8, o);
#pragma unroll
      for (int i = 0; i < 8; i++) q[i] = __shfl_sync(FULL, r1, 4 * i);
      q[8] = q8;
    }
    LSPoint pt[3];
#pragma unroll
    for (int a = 0; a < 3; a++) {
      // No FMA contraction: with fused multiply-adds the derivative at a Newton iterate no longer rounds to the
      // value that ends MJX's bracketing and the loop runs to ls_iterations (measured 100 vs 28 line-search
      // iterations per substep).  Mirrors the oracle's unfused arithmetic.
      const float t0 = q[3 * a] + qg0, t1 = q[3 * a + 1] + qg1, t2 = q[3 * a + 2] + qg2;
      pt[a].alpha = al[a];
      pt[a].cost = __fadd_rn(__fadd_rn(__fmul_rn(__fmul_rn(al[a], al[a]), t2), __fmul_rn(al[a], t1)), t0);
      pt[a].d0 = __fadd_rn(__fmul_rn(__fmul_rn(2.f, al[a]), t2), t1);
      pt[a].d1 = __fadd_rn(__fmul_rn(2.f, t2), (t2 == 0.f ? MJ_MINVAL : 0.f));
    }
    // ---- bracketing logic
    if (phase == 0) {
      p0 = pt[0];
      al[0] = al[1] = al[2] = p0.alpha - p0.d0 / p0.d1;
      phase = 1;
      continue;
    }
    bool swap = true;
    if (phase == 1) {
      const LSPoint lo0 = pt[0];
      const bool lesser = lo0.d0 < p0.d0;
      hi = lesser ? p0 : lo0;
      lo = lesser ? lo0 : p0;
      phase = 2;
    } else {
      const LSPoint lo_next = pt[0], hi_next = pt[1], mid = pt[2];
      const bool swap_lo_next = (lo.d0 > 0.f) || (lo.d0 < lo_next.d0);
      if (swap_lo_next) lo = lo_next;
      const bool swap_lo_mid = (mid.d0 < 0.f) && (lo.d0 < mid.d0);
      if (swap_lo_mid) lo = mid;
      const bool swap_hi_next = (hi.d0 < 0.f) || (hi.d0 > hi_next.d0);
      if (swap_hi_next) hi = hi_next;
      const bool swap_hi_mid = (mid.d0 > 0.f) && (hi.d0 > mid.d0);
      if (swap_hi_mid) hi = mid;
      swap = swap_lo_next || swap_lo_mid || swap_hi_next || swap_hi_mid;
      it++;
      // MJX's bracketing often falls into an exact cycle (lo and hi trade places across a kink of the piecewise-
      // linear derivative, or hop between neighbouring floats; periods 2, 3 and 4 all occur) and then runs to
      // ls_iterations.  A bracket is a pure function of (lo.alpha, hi.alpha), so once the pair equals an anchor taken
      // at a power-of-two iteration (Brent) every later state is known: skip whole periods and run only the remainder.
      // Bit-identical result, 11.3 -> 3.3 evaluated iterations per line search (profiles/README.md).
      if (swap && !jumped && it < dm->ls_iterations) {
        if (lo.alpha == anchor_lo && hi.alpha == anchor_hi) {
          it = dm->ls_iterations - (dm->ls_iterations - it) % (it - anchor_it);
          jumped = true;
        } else if ((it & (it - 1)) == 0) {
          anchor_lo = lo.alpha; anchor_hi = hi.alpha; anchor_it = it;
        }
      }
    }
    bool done = it >= dm->ls_iterations;
    done |= !swap;
    done |= (lo.d0 < 0.f) && (lo.d0 > -gtol);
    done |= (hi.d0 > 0.f) && (hi.d0 < gtol);
    if (done) break;
    al[0] = lo.alpha - lo.d0 / lo.d1;
    al[1] = hi.alpha - hi.d0 / hi.d1;
    al[2] = 0.5f * (lo.alpha + hi.alpha);
  }
  const bool improved = (lo.cost < p0.cost) || (hi.cost < p0.cost);
  const float alpha = lo.cost < hi.cost ? lo.alpha : hi.alpha;
  if (improved) {
    if (lane < nv) {
      sm[ar::V_QACC + lane] += sm[ar::V_SEARCH + lane] * alpha;
      sm[ar::V_MA + lane] += sm[ar::V_MV + lane] * alpha;
    }
#pragma unroll 1
    for (int r = lane; r < nrow; r += 32) sm[ar::E_JAREF + r] += sm[ar::E_JV + r] * alpha;
  }
  RSRX_SYNC();
  return it;
}

// solver.py::solve.  Returns niter | (total line-search iterations << 8).
__device__ __noinline__ int solve(const DModel* __restrict__ dm, float* sm, int lane, int nsr, int ncon, int* status) {
  const int nv = dm->nv, nrow = nsr + 6 * ncon;
  if (nrow == 0) {
    if (lane < nv) { sm[ar::V_QACC + lane] = sm[ar::V_QACCS + lane]; sm[ar::V_QFRCC + lane] = 0.f; }
    RSRX_SYNC();
    return 0;
  }
  float gauss;
  // warm start: whichever of qacc_warmstart / qacc_smooth costs less.  The warm start usually wins, so it is
  // evaluated last and its context is simply kept.
  const float cs = ctx_create(dm, sm, lane, nsr, ncon, sm + ar::V_QACCS, &gauss);
  const float cw = ctx_create(dm, sm, lane, nsr, ncon, sm + ar::WARM, &gauss);
  float cost = cw;
  if (!(cw < cs)) cost = ctx_create(dm, sm, lane, nsr, ncon, sm + ar::V_QACCS, &gauss);
  float prev_cost = INFINITY;
  const float scale = 1.f / (dm->meaninertia * (float)(nv > 1 ? nv : 1));
  int niter = 0, ls_total = 0;
  bool changed = true, have_factor = false;
#pragma unroll 1
  for (;;) {
    // _update_gradient, split: the gradient first, on which (with the cost improvement) MJX's loop condition is
    // tested; the Newton direction (H assembly, Cholesky, solve) only if another iteration follows.  MJX computes it
    // in the last pass too and throws it away: same qacc, one factorisation + solve less per mjx.step.
    float g = 0.f;
    if (lane < nv) {
      g = sm[ar::V_MA + lane] - sm[ar::V_SMOOTH + lane] - sm[ar::V_QFRCC + lane];
      sm[ar::V_GRAD + lane] = g;
      sm[ar::V_MGRAD + lane] = g;
      g *= g;
    }
    const float improvement = (prev_cost - cost) * scale;
    const float gradient = sqrtf(warp_sum(g)) * scale;
    bool done = niter >= dm->iterations;
    done |= improvement < dm->tolerance;
    done |= gradient < dm->tolerance;
    // a non-finite state (an exploded simulation; MJX propagates NaN the same way) cannot satisfy either test and would
    // spin through iterations x ls_iterations passes for a result that is NaN anyway, stalling its whole CTA
    done |= !(cost == cost) || !(gradient == gradient);
    if (done) break;
    RSRX_SYNC();
    update_gradient(dm, sm, lane, nsr, ncon, !changed && have_factor);
    have_factor = true;
    if (lane < nv) sm[ar::V_SEARCH + lane] = -sm[ar::V_MGRAD + lane];
    RSRX_SYNC();
    ls_total += linesearch(dm, sm, lane, nsr, ncon, gauss);
    prev_cost = cost;
    cost = update_constraint(dm, sm, lane, nsr, ncon, &gauss, &changed);
    niter++;
  }
  if (niter >= dm->iterations) *status |= RSRX_STATUS_SOLVER_CAP;
  if (lane < nv) sm[ar::WARM + lane] = sm[ar::V_QACC + lane];
  RSRX_SYNC();
  return niter | (ls_total << 8);
}

// Phase barriers (step_kernel only).  The kernel is bound by instruction delivery: the L1 instruction
// cache hit rate is 60 % when the CTA's warps sit in different phases, because the one-shot phases stream
// ~130 KB of code per substep through it and evict the solver's loops.  Re-aligning the warps of a CTA at
// phase boundaries lets one fetch serve all of them.  RSRX_SYNC_MASK selects the boundaries (bit 0: substep
// start, 1: collision, 2: make_constraint, 3: velocity/forces, 4: solve, 5: integrate).
#ifndef RSRX_SYNC_MASK
#define RSRX_SYNC_MASK 37  // substep start + before make_constraint + after the solver.  Round 1 timed the first steps after
                          // reset, where start + after-solver (33) was best (2.37 vs 2.51 ms with the start barrier alone);
                          // on the stationary episode-phase distribution bench.py times since round 2 the contact load and
                          // its spread across the 19 envs of a CTA are larger, the instruction-cache hit rate drops to 79 %,
                          // and a third alignment point pays: 2.41 vs 2.50 ms at 8192 envs, 3.20 vs 3.30 on the T-shape env,
                          // equal or better down to 128 envs (profiles/README.md; every 3+-barrier mask lands at 2.40-2.42,
                          // no barrier at all: 5.9 ms)
#endif
#ifndef RSRX_BAR_GROUPS
#define RSRX_BAR_GROUPS 1
#endif
constexpr int kPhaseBarriers = __builtin_popcount(RSRX_SYNC_MASK & 0x3f);
template <bool SYNC>
__device__ __forceinline__ void phase_barrier(int bit) {
  if (SYNC && ((RSRX_SYNC_MASK >> bit) & 1)) {
#if RSRX_BAR_GROUPS == 1
    __syncthreads();
#else
    // experiment: the CTA's warps re-align in RSRX_BAR_GROUPS independent groups (named barriers 1..)
    constexpr int per = RSRX_WPB / RSRX_BAR_GROUPS;
    asm volatile("bar.sync %0, %1;" ::"r"(1 + (int)(threadIdx.x >> 5) / per), "n"(32 * per) : "memory");
#endif
  }
}

// forward.py::forward.  Returns niter; fills dims.
template <bool SYNC>
__device__ __forceinline__ int forward(const DModel* __restrict__ dm, float* sm, int lane, SolverDims* sd, int* status) {
  phase_barrier<SYNC>(0);
  kinematics(dm, sm, lane);
  com_pos(dm, sm, lane);
  crb_and_factor(dm, sm, lane);
  phase_barrier<SYNC>(3);
  velocity_and_forces(dm, sm, lane);  // before collision: its scratch (region P) is reused by the constraint rows
  phase_barrier<SYNC>(1);
  const int ncon = collision(dm, sm, lane, status);
  phase_barrier<SYNC>(2);
  const int nsr = make_constraint(dm, sm, lane, ncon);
  sd->nsr = nsr; sd->ncon = ncon; sd->nrow = nsr + 6 * ncon;
  phase_barrier<SYNC>(4);
  const int r = solve(dm, sm, lane, nsr, ncon, status);
  phase_barrier<SYNC>(5);
  return r;
}

// forward.py::implicit + _advance (implicitfast; only dof damping contributes to qDeriv)
__device__ __noinline__ void implicit_advance(const DModel* __restrict__ dm, float* sm, int lane) {
  const int nv = dm->nv;
  const float dt = dm->timestep;
  if (lane < nv) sm[ar::V_TMP + lane] = sm[ar::V_SMOOTH + lane] + sm[ar::V_QFRCC + lane];
  copy_M_permuted(dm, sm, lane, sm + ar::DAMP, dt);
  warp_chol_factor(dm, sm, lane, true);
  warp_chol_solve(dm, sm, sm + ar::V_TMP, lane, true);
  if (lane < nv) sm[ar::QVEL + lane] += sm[ar::V_TMP + lane] * dt;
  RSRX_SYNC();
  if (lane < dm->njnt) {
    const int j = lane, qa = dm->jnt_qposadr[j], d = dm->jnt_dofadr[j];
    if (dm->jnt_type[j] == RSRX_JNT_FREE) {
      for (int i = 0; i < 3; i++) sm[ar::QPOS + qa + i] += dt * sm[ar::QVEL + d + i];
      float v[3] = {sm[ar::QVEL + d + 3], sm[ar::QVEL + d + 4], sm[ar::QVEL + d + 5]};
      const float nrm = sqrtf(dot3(v, v));
      if (nrm > 0.f) { v[0] /= nrm; v[1] /= nrm; v[2] /= nrm; }
      float s, c;
      sincosf(dt * nrm * 0.5f, &s, &c);
      float qr[4] = {c, v[0] * s, v[1] * s, v[2] * s}, q2[4];
      float q0[4] = {sm[ar::QPOS + qa + 3], sm[ar::QPOS + qa + 4], sm[ar::QPOS + qa + 5], sm[ar::QPOS + qa + 6]};
      quat_mul(q2, q0, qr);
      normalize4(q2);
      for (int i = 0; i < 4; i++) sm[ar::QPOS + qa + 3 + i] = q2[i];
    } else {
      sm[ar::QPOS + qa] += dt * sm[ar::QVEL + d];
    }
  }
  RSRX_SYNC();
}

}  // namespace RSRX_NS
