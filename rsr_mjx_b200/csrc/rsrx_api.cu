// rsrx_api.cu — C-ABI of librsrx.so (include/rsrx.h): model upload and kernel
// launches.  No torch types; the caller owns every state buffer.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "rsrx_env.cuh"
#include "rsrx_loss.cuh"
#include "rsrx_ppo.cuh"
#include "rsrx_gemm.cuh"
#include "rsrx_mid.h"
#include "rsrx_mlp.cuh"
#include "rsrx_redo.h"

using namespace rsrx;

struct rsrx_model {
  DModel host;
  DModel* dev;
  int smem_bytes;   // dynamic shared memory of a full CTA (WPB envs)
  int num_sms;
  // overflow of the shared-memory Jacobian-row pool: [spill_envs][ar::SPILL_STRIDE] floats, grown on demand (the only
  // allocation reset/step ever make: the first call for a larger N, never inside a stream capture)
  float* spill = nullptr;
  int spill_envs = 0;
  // envs whose substep exceeds the fast arena's contact capacity are re-run by the large-capacity instantiation
  // (rsrx_redo.cu): its device model, launch shape, and the device list the fast kernels append to ([0] count,
  // [1] ticket, [2..] env ids; spill_envs entries)
  void* big_dev = nullptr;
  int big_smem = 0, big_grid = 0;
  // the mid-capacity step kernel (rsrx_mid.cu: 64 contacts, <= 8 envs per CTA) takes the step launch of small batches
  // (N <= num_sms * 8), where shared memory is plentiful, so that the redo pass is practically never needed there
  void* mid_dev = nullptr;
  int mid_arena_bytes = 0;
  int* redo = nullptr;
  // buffers outgrown by a larger batch: launches already queued may still use them, so they are kept until the model is
  // destroyed instead of synchronising the device inside reset / step
  std::vector<void*> retired;
};

// Launch shape for N envs: one CTA per SM per round, the rounds as evenly filled as possible.  8192 envs on 148 SMs:
// 3 rounds of 19 envs per CTA (432 CTAs); 1024 envs: one round of 147 CTAs x 7 envs instead of 54 CTAs x 19 on a third
// of the SMs.
constexpr int kSmallW = WPB < 14 ? WPB : 14;  // CTAs of up to 14 warps use the 128-register instantiation of step_kernel
struct LaunchCfg { int grid, block; size_t smem; };
static LaunchCfg launch_cfg(const rsrx_model* m, int N) {
  const int per_round = m->num_sms * WPB;
  const int rounds = (N + per_round - 1) / per_round;
  int w = (N + m->num_sms * rounds - 1) / (m->num_sms * rounds);
  static const int forced = getenv("RSRX_FORCE_WPB") ? atoi(getenv("RSRX_FORCE_WPB")) : 0;  // experiments only
  if (forced > 0) w = forced;
  w = w < 1 ? 1 : (w > WPB ? WPB : w);
  return {(N + w - 1) / w, 32 * w, (size_t)w * m->host.arena_stride * sizeof(float)};
}

static thread_local std::string g_err;
static int fail(const std::string& s) { g_err = s; return 1; }
#define CUDA_OK(x)                                                                                   \
  do {                                                                                               \
    cudaError_t _e = (x);                                                                            \
    if (_e != cudaSuccess) return fail(std::string(#x) + ": " + cudaGetErrorString(_e));             \
  } while (0)

extern "C" const char* rsrx_last_error(void) { return g_err.c_str(); }
extern "C" const char* rsrx_version(void) { return "rsrx 0.1 (sm_100a)"; }
extern "C" size_t rsrx_model_blob_size(void) { return sizeof(rsrx_model_blob); }
extern "C" size_t rsrx_env_cfg_size(void) { return sizeof(rsrx_env_cfg); }
extern "C" int rsrx_debug_stride(void) { return dbg::STRIDE; }
extern "C" int rsrx_max_contacts(void) { return MAXC; }

// ---- host-side double math for the static precomputation -----------------------
namespace {
void hq_mul(double* r, const double* u, const double* v) {
  double a = u[0] * v[0] - u[1] * v[1] - u[2] * v[2] - u[3] * v[3];
  double b = u[0] * v[1] + u[1] * v[0] + u[2] * v[3] - u[3] * v[2];
  double c = u[0] * v[2] - u[1] * v[3] + u[2] * v[0] + u[3] * v[1];
  double d = u[0] * v[3] + u[1] * v[2] - u[2] * v[1] + u[3] * v[0];
  r[0] = a; r[1] = b; r[2] = c; r[3] = d;
}
void hq_mat(double* m, const double* q) {
  m[0] = q[0] * q[0] + q[1] * q[1] - q[2] * q[2] - q[3] * q[3]; m[1] = 2 * (q[1] * q[2] - q[0] * q[3]); m[2] = 2 * (q[1] * q[3] + q[0] * q[2]);
  m[3] = 2 * (q[1] * q[2] + q[0] * q[3]); m[4] = q[0] * q[0] - q[1] * q[1] + q[2] * q[2] - q[3] * q[3]; m[5] = 2 * (q[2] * q[3] - q[0] * q[1]);
  m[6] = 2 * (q[1] * q[3] - q[0] * q[2]); m[7] = 2 * (q[2] * q[3] + q[0] * q[1]); m[8] = q[0] * q[0] - q[1] * q[1] - q[2] * q[2] + q[3] * q[3];
}
void hq_rot(double* r, const double* v, const double* q) {
  double m[9];
  hq_mat(m, q);
  double x = m[0] * v[0] + m[1] * v[1] + m[2] * v[2], y = m[3] * v[0] + m[4] * v[1] + m[5] * v[2], z = m[6] * v[0] + m[7] * v[1] + m[8] * v[2];
  r[0] = x; r[1] = y; r[2] = z;
}
}  // namespace

static int build_dmodel(const rsrx_model_blob& b, const rsrx_env_cfg& c, DModel& d) {
  memset(&d, 0, sizeof(d));
  if (b.magic != RSRX_MAGIC || b.version != RSRX_VERSION) return fail("model blob: bad magic/version");
  if (b.nbody > NB || b.njnt > NJ || b.nq > NQ || b.nv > NV || b.nu > NU || b.ngeom > NG || b.nsite > NS || b.npair > NP ||
      b.neq > RSRX_MAXEQ)
    return fail("model blob exceeds compiled capacities");
  if (b.neq + b.nv > 32 || b.njnt > 32) return fail("too many constraint candidates for one warp");
  {
    int nsr = b.neq;
    for (int i = 0; i < b.nv; i++) nsr += b.dof_frictionloss[i] > 0;
    for (int j = 0; j < b.njnt; j++) nsr += (b.jnt_limited[j] && b.jnt_type[j] != RSRX_JNT_FREE);
    if (nsr > MAXSR) return fail("too many equality / friction / limit rows");
  }
  d.nbody = b.nbody; d.njnt = b.njnt; d.nq = b.nq; d.nv = b.nv; d.nu = b.nu; d.ngeom = b.ngeom; d.nsite = b.nsite;
  d.npair = b.npair; d.neq = b.neq;
  d.iterations = b.iterations; d.ls_iterations = b.ls_iterations;
  d.timestep = (float)b.timestep;
  for (int i = 0; i < 3; i++) d.gravity[i] = (float)b.gravity[i];
  d.tolerance = (float)b.tolerance; d.ls_tolerance = (float)b.ls_tolerance; d.impratio = (float)b.impratio;
  d.meaninertia = (float)b.meaninertia;
  int maxdepth = 0;
  double sx[NB][3] = {{0}}, sq[NB][4] = {{0}};
  sq[0][0] = 1;
  for (int i = 0; i < b.nbody; i++) {
    d.body_parentid[i] = b.body_parentid[i]; d.body_rootid[i] = b.body_rootid[i]; d.body_jntadr[i] = b.body_jntadr[i];
    d.body_jntnum[i] = b.body_jntnum[i]; d.body_dofadr[i] = b.body_dofadr[i]; d.body_dofnum[i] = b.body_dofnum[i];
    d.body_depth[i] = b.body_depth[i];
    d.body_static[i] = (b.body_weldid[i] == 0);
    if (b.body_depth[i] > maxdepth) maxdepth = b.body_depth[i];
    for (int k = 0; k < 3; k++) { d.body_pos[i][k] = (float)b.body_pos[i][k]; d.body_ipos[i][k] = (float)b.body_ipos[i][k]; d.body_inertia[i][k] = (float)b.body_inertia[i][k]; }
    for (int k = 0; k < 4; k++) { d.body_quat[i][k] = (float)b.body_quat[i][k]; d.body_iquat[i][k] = (float)b.body_iquat[i][k]; }
    d.body_mass[i] = (float)b.body_mass[i];
    d.body_invweight0[i] = (float)b.body_invweight0[i][0];
    if (i > 0 && i <= b.body_parentid[i]) return fail("bodies must be in depth-first order");
    // subtree end (bodies are in DFS pre-order)
    int end = i + 1;
    while (end < b.nbody && b.body_depth[end] > b.body_depth[i]) end++;
    d.body_subtree_end[i] = end;
    // dof mask along the path to the world
    uint32_t mask = 0;
    for (int a = i; a > 0; a = b.body_parentid[a])
      for (int k = 0; k < b.body_dofnum[a]; k++) mask |= 1u << (b.body_dofadr[a] + k);
    d.body_dofmask[i] = mask;
    // static world pose
    if (i > 0 && d.body_static[i]) {
      int p = b.body_parentid[i];
      double t[3];
      hq_rot(t, b.body_pos[i], sq[p]);
      for (int k = 0; k < 3; k++) sx[i][k] = sx[p][k] + t[k];
      hq_mul(sq[i], sq[p], b.body_quat[i]);
    }
    for (int k = 0; k < 3; k++) d.static_xpos[i][k] = (float)sx[i][k];
    for (int k = 0; k < 4; k++) d.static_xquat[i][k] = (float)sq[i][k];
  }
  d.nlevel = maxdepth + 1;
  d.ntree = 0;
  for (int i = 0; i < b.nbody; i++) d.body_treeid[i] = 0;
  for (int i = 1; i < b.nbody; i++) {
    if (b.body_parentid[i] == 0) {
      if (d.ntree >= MAXTREE) return fail("too many kinematic trees");
      d.tree_dofadr[d.ntree] = 0; d.tree_dofnum[d.ntree] = 0;
      d.body_treeid[i] = d.ntree++;
    } else {
      d.body_treeid[i] = d.body_treeid[b.body_rootid[i]];
    }
  }
  for (int t = 0; t < d.ntree; t++) {
    int lo = 1 << 30, hi = -1;
    for (int v = 0; v < b.nv; v++)
      if (d.body_treeid[b.dof_bodyid[v]] == t) { if (v < lo) lo = v; if (v > hi) hi = v; }
    if (hi >= 0) {
      d.tree_dofadr[t] = lo; d.tree_dofnum[t] = hi - lo + 1;
      for (int v = lo; v <= hi; v++) if (d.body_treeid[b.dof_bodyid[v]] != t) return fail("dofs of a kinematic tree must be contiguous");
      for (int v = lo; v <= hi; v++) { d.dof_tree_lo[v] = lo; d.dof_tree_hi[v] = hi; }
    }
  }
  for (int j = 0; j < b.njnt; j++) {
    d.jnt_type[j] = b.jnt_type[j]; d.jnt_qposadr[j] = b.jnt_qposadr[j]; d.jnt_dofadr[j] = b.jnt_dofadr[j];
    d.jnt_bodyid[j] = b.jnt_bodyid[j]; d.jnt_limited[j] = b.jnt_limited[j];
    if (b.jnt_type[j] != RSRX_JNT_FREE && b.jnt_type[j] != RSRX_JNT_HINGE && b.jnt_type[j] != RSRX_JNT_SLIDE)
      return fail("unsupported joint type");
    for (int k = 0; k < 3; k++) { d.jnt_pos[j][k] = (float)b.jnt_pos[j][k]; d.jnt_axis[j][k] = (float)b.jnt_axis[j][k]; }
    for (int k = 0; k < 2; k++) { d.jnt_range[j][k] = (float)b.jnt_range[j][k]; d.jnt_solref[j][k] = (float)b.jnt_solref[j][k]; }
    for (int k = 0; k < 5; k++) d.jnt_solimp[j][k] = (float)b.jnt_solimp[j][k];
    d.jnt_margin[j] = (float)b.jnt_margin[j];
  }
  for (int i = 0; i < b.nq; i++) d.qpos0[i] = (float)b.qpos0[i];
  d.nment = 0;
  for (int i = 0; i < b.nv; i++) {
    d.dof_bodyid[i] = b.dof_bodyid[i]; d.dof_jntid[i] = b.dof_jntid[i]; d.dof_parentid[i] = b.dof_parentid[i];
    int j = b.dof_jntid[i];
    d.dof_actfrclimited[i] = b.jnt_actfrclimited[j];
    d.dof_actfrcrange[i][0] = (float)b.jnt_actfrcrange[j][0]; d.dof_actfrcrange[i][1] = (float)b.jnt_actfrcrange[j][1];
    d.dof_hasfriction[i] = b.dof_frictionloss[i] > 0;
    d.dof_damping[i] = (float)b.dof_damping[i]; d.dof_frictionloss[i] = (float)b.dof_frictionloss[i];
    d.dof_armature[i] = (float)b.dof_armature[i]; d.dof_invweight0[i] = (float)b.dof_invweight0[i];
    for (int k = 0; k < 2; k++) d.dof_solref[i][k] = (float)b.dof_solref[i][k];
    for (int k = 0; k < 5; k++) d.dof_solimp[i][k] = (float)b.dof_solimp[i][k];
    for (int a = i; a >= 0; a = b.dof_parentid[a]) {
      if (d.nment >= MAXMENT) return fail("mass-matrix entry list overflow");
      d.ment_i[d.nment] = (unsigned char)i; d.ment_j[d.nment] = (unsigned char)a; d.nment++;
    }
  }
  d.ntri = 0;
  for (int i = 0; i < b.nv; i++)
    for (int j = 0; j <= i; j++) { d.tri_i[d.ntri] = (unsigned char)i; d.tri_j[d.ntri] = (unsigned char)j; d.ntri++; }
  for (int g = 0; g < b.ngeom; g++) {
    d.geom_type[g] = b.geom_type[g]; d.geom_bodyid[g] = b.geom_bodyid[g];
    int bd = b.geom_bodyid[g];
    d.geom_static[g] = d.body_static[bd];
    for (int k = 0; k < 3; k++) { d.geom_pos[g][k] = (float)b.geom_pos[g][k]; d.geom_size[g][k] = (float)b.geom_size[g][k]; d.geom_friction[g][k] = (float)b.geom_friction[g][k]; }
    for (int k = 0; k < 4; k++) d.geom_quat[g][k] = (float)b.geom_quat[g][k];
    d.geom_rbound[g] = std::sqrt(d.geom_size[g][0] * d.geom_size[g][0] + d.geom_size[g][1] * d.geom_size[g][1] + d.geom_size[g][2] * d.geom_size[g][2]);
    if (d.geom_static[g]) {
      double t[3], q[4], m[9];
      hq_rot(t, b.geom_pos[g], sq[bd]);
      hq_mul(q, sq[bd], b.geom_quat[g]);
      hq_mat(m, q);
      for (int k = 0; k < 3; k++) d.geom_static_xpos[g][k] = (float)(sx[bd][k] + t[k]);
      for (int k = 0; k < 9; k++) d.geom_static_xmat[g][k] = (float)m[k];
    }
  }
  for (int s = 0; s < b.nsite; s++) {
    d.site_bodyid[s] = b.site_bodyid[s];
    for (int k = 0; k < 3; k++) d.site_pos[s][k] = (float)b.site_pos[s][k];
  }
  for (int p = 0; p < b.npair; p++) {
    int g1 = b.pair_geom1[p], g2 = b.pair_geom2[p];
    d.pair_g1[p] = g1; d.pair_g2[p] = g2;
    int t1 = b.geom_type[g1], t2 = b.geom_type[g2];
    if (!((t1 == RSRX_GEOM_PLANE || t1 == RSRX_GEOM_BOX) && t2 == RSRX_GEOM_BOX)) return fail("only plane-box and box-box pairs are supported");
    int condim = b.geom_condim[g1] > b.geom_condim[g2] ? b.geom_condim[g1] : b.geom_condim[g2];
    if (condim != 4) return fail("only condim 4 contacts are supported (all Airbot pairs are condim 4)");
    // float32 mixing, like MJX does at run time
    float mix1 = (float)b.geom_solmix[g1], mix2 = (float)b.geom_solmix[g2];
    float mix = mix1 / (mix1 + mix2);
    if (mix1 < MJ_MINVAL && mix2 < MJ_MINVAL) mix = 0.5f;
    else if (mix1 < MJ_MINVAL) mix = 0.f;
    else if (mix2 < MJ_MINVAL) mix = 1.f;
    float r1[2] = {(float)b.geom_solref[g1][0], (float)b.geom_solref[g1][1]}, r2[2] = {(float)b.geom_solref[g2][0], (float)b.geom_solref[g2][1]};
    bool standard = r1[0] > 0 && r2[0] > 0;
    for (int k = 0; k < 2; k++) d.pair_solref[p][k] = standard ? mix * r1[k] + (1.f - mix) * r2[k] : (r1[k] < r2[k] ? r1[k] : r2[k]);
    for (int k = 0; k < 5; k++) d.pair_solimp[p][k] = mix * (float)b.geom_solimp[g1][k] + (1.f - mix) * (float)b.geom_solimp[g2][k];
    d.pair_margin[p] = (float)(b.geom_margin[g1] > b.geom_margin[g2] ? b.geom_margin[g1] : b.geom_margin[g2]);
    d.pair_tran[p] = (float)b.body_invweight0[b.geom_bodyid[g1]][0] + (float)b.body_invweight0[b.geom_bodyid[g2]][0];
    {
      int b1 = b.geom_bodyid[g1], b2 = b.geom_bodyid[g2];
      int na = d.body_dofmask[b1] ? d.tree_dofnum[d.body_treeid[b1]] : 0, nb = d.body_dofmask[b2] ? d.tree_dofnum[d.body_treeid[b2]] : 0;
      if (na && nb && d.body_treeid[b1] == d.body_treeid[b2]) return fail("collision pairs inside one kinematic tree are not supported");
      if (na + nb > NCOL) return fail("a collision pair couples more dofs than the kernel's Jacobian width");
    }
  }
  // ---- dof blocks: dofs of one kinematic tree, plus trees joined by a collision pair, form one block of H
  {
    int comp[NV];
    for (int i = 0; i < b.nv; i++) comp[i] = b.body_rootid[b.dof_bodyid[i]];
    auto relabel = [&](int from, int to) { for (int i = 0; i < b.nv; i++) if (comp[i] == from) comp[i] = to; };
    bool nz[NV][NV] = {{false}};
    for (int e = 0; e < d.nment; e++) nz[d.ment_i[e]][d.ment_j[e]] = true;
    for (int p = 0; p < b.npair; p++) {
      int b1 = b.geom_bodyid[b.pair_geom1[p]], b2 = b.geom_bodyid[b.pair_geom2[p]];
      uint32_t mask = d.body_dofmask[b1] | d.body_dofmask[b2];
      for (int i = 0; i < b.nv; i++)
        for (int j = 0; j <= i; j++)
          if (((mask >> i) & 1u) && ((mask >> j) & 1u)) nz[i][j] = true;
      if (d.body_dofmask[b1] && d.body_dofmask[b2]) {
        int c1 = -1, c2 = -1;
        for (int i = 0; i < b.nv; i++) { if ((d.body_dofmask[b1] >> i) & 1u) c1 = comp[i]; if ((d.body_dofmask[b2] >> i) & 1u) c2 = comp[i]; }
        if (c1 != c2) relabel(c2, c1);
      }
    }
    int pos = 0;
    bool placed[NV] = {false};
    for (int i = 0; i < b.nv; i++) {
      if (placed[i]) continue;
      int start = pos;
      for (int j = i; j < b.nv; j++)
        if (!placed[j] && comp[j] == comp[i]) { placed[j] = true; d.pos_of_dof[j] = pos; d.dof_of_pos[pos] = j; pos++; }
      for (int q = start; q < pos; q++) { d.blk_start[q] = start; d.blk_end[q] = pos - 1; }
    }
    for (int e = 0; e < d.ntri; e++) {
      int i = d.tri_i[e], j = d.tri_j[e], pi = d.pos_of_dof[i], pj = d.pos_of_dof[j];
      d.tri_src[e] = (unsigned short)(i * LD + j);
      d.tri_dst[e] = (unsigned short)(tri(pi > pj ? pi : pj) + (pi > pj ? pj : pi));
    }
    for (int q = 0; q < b.nv; q++) {  // per-tree sub-blocks inside the permuted order
      int t = d.body_treeid[b.dof_bodyid[d.dof_of_pos[q]]], lo = q, hi = q;
      while (lo > 0 && d.body_treeid[b.dof_bodyid[d.dof_of_pos[lo - 1]]] == t) lo--;
      while (hi + 1 < b.nv && d.body_treeid[b.dof_bodyid[d.dof_of_pos[hi + 1]]] == t) hi++;
      d.tblk_start[q] = lo; d.tblk_end[q] = hi;
    }
    d.blk_max = d.tblk_max = 1;
    for (int q = 0; q < b.nv; q++) {
      d.blk_max = std::max(d.blk_max, d.blk_end[q] - d.blk_start[q] + 1);
      d.tblk_max = std::max(d.tblk_max, d.tblk_end[q] - d.tblk_start[q] + 1);
    }
    d.nhent = 0;
    for (int i = 0; i < b.nv; i++)
      for (int j = 0; j <= i; j++)
        if (nz[i][j]) {
          if (d.blk_start[d.pos_of_dof[i]] != d.blk_start[d.pos_of_dof[j]]) return fail("internal: H entry outside its block");
          d.hent_i[d.nhent] = (unsigned char)i; d.hent_j[d.nhent] = (unsigned char)j; d.nhent++;
        }
  }
  for (int u = 0; u < b.nu; u++) {
    int j = b.act_trnid[u];
    d.act_qadr[u] = b.jnt_qposadr[j]; d.act_dof[u] = b.jnt_dofadr[j];
    d.act_ctrllimited[u] = b.act_ctrllimited[u]; d.act_forcelimited[u] = b.act_forcelimited[u];
    d.act_gear[u] = (float)b.act_gear[u]; d.act_gain[u] = (float)b.act_gainprm[u][0];
    for (int k = 0; k < 3; k++) d.act_bias[u][k] = (float)b.act_biasprm[u][k];
    for (int k = 0; k < 2; k++) { d.act_ctrlrange[u][k] = (float)b.act_ctrlrange[u][k]; d.act_forcerange[u][k] = (float)b.act_forcerange[u][k]; }
  }
  for (int e = 0; e < b.neq; e++) {
    int j1 = b.eq_obj1id[e], j2 = b.eq_obj2id[e];
    d.eq_q1[e] = b.jnt_qposadr[j1]; d.eq_d1[e] = b.jnt_dofadr[j1];
    d.eq_q2[e] = j2 >= 0 ? b.jnt_qposadr[j2] : -1; d.eq_d2[e] = j2 >= 0 ? b.jnt_dofadr[j2] : -1;
    float invw = (float)b.dof_invweight0[b.jnt_dofadr[j1]];
    if (j2 >= 0) invw += (float)b.dof_invweight0[b.jnt_dofadr[j2]];
    d.eq_invweight[e] = invw;
    for (int k = 0; k < 5; k++) { d.eq_data[e][k] = (float)b.eq_data[e][k]; d.eq_solimp[e][k] = (float)b.eq_solimp[e][k]; }
    for (int k = 0; k < 2; k++) d.eq_solref[e][k] = (float)b.eq_solref[e][k];
  }
  // env
  d.env_kind = c.env_kind; d.episode_length = c.episode_length; d.action_repeat = c.action_repeat; d.n_frames = c.n_frames;
  d.cube_body = c.cube_body; d.target_body = c.target_body; d.site_endpoint = c.site_endpoint; d.site_tail = c.site_tail;
  d.site_target_tail = c.site_target_tail; d.geom_base = c.geom_base; d.geom_vertical = c.geom_vertical;
  d.geom_target_base = c.geom_target_base; d.geom_target_vertical = c.geom_target_vertical;
  for (int i = 0; i < 6; i++) d.joint_qadr[i] = c.joint_qadr[i];
  for (int i = 0; i < NU; i++) d.action_scale[i] = (float)c.action_scale[i];
  d.push_reward_weight = (float)c.push_reward_weight; d.siet_to_box_reward_weight = (float)c.siet_to_box_reward_weight;
  d.healthy_reward = (float)c.healthy_reward; d.endpoint_min_z_pos = (float)c.endpoint_min_z_pos;
  if (c.env_kind < 0 || c.env_kind > 2) return fail("bad env_kind");
  if (b.nu < 5) return fail("Airbot envs need 5 actuators");
  // layout
  rsrx_layout& L = d.lay;
  int o = 0;
  L.qpos = o; o += b.nq;
  L.qvel = o; o += b.nv;
  L.ctrl = o; o += b.nu;
  L.qacc_warmstart = o; o += b.nv;
  L.time = o; o += 1;
  L.xpos = o; o += b.nbody * 3;
  L.xquat = o; o += b.nbody * 4;
  L.site_xpos = o; o += b.nsite * 3;
  L.geom_xpos = o; o += b.ngeom * 3;
  L.data_stride = (o + 3) / 4 * 4;
  L.obs_size = c.env_kind == RSRX_ENV_T ? 16 : 23;
  L.obs_stride = OBS_STRIDE;
  L.info_stride = RSRX_INFO_STRIDE;
  L.metrics_stride = METRICS_STRIDE;
  L.nq = b.nq; L.nv = b.nv; L.nu = b.nu; L.nbody = b.nbody; L.nsite = b.nsite; L.ngeom = b.ngeom;
  return 0;
}

extern "C" int rsrx_model_create(const void* blob_host, size_t blob_bytes, const rsrx_env_cfg* cfg_host, rsrx_model** out) {
  if (!blob_host || !cfg_host || !out) return fail("rsrx_model_create: null argument");
  if (blob_bytes != sizeof(rsrx_model_blob)) return fail("rsrx_model_create: blob size mismatch (host/lib out of sync)");
  rsrx_model* m = new rsrx_model();
  if (build_dmodel(*reinterpret_cast<const rsrx_model_blob*>(blob_host), *cfg_host, m->host)) { delete m; return 1; }
  {
    // the arena fills the SM's shared memory at WPB envs per CTA: whatever is left after the fixed part is the
    // Jacobian-row pool (B200: 232448 B / 19 envs = 3058 floats per env, 2176 fixed, 882 pool)
    int dev = 0, max_smem = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&m->num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || m->num_sms <= 0)
      m->num_sms = 148;
    if (cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || max_smem <= 0)
      max_smem = 227 * 1024;
    int stride = (max_smem / (int)sizeof(float) / WPB) & ~3;  // 16-byte multiples: the arena starts with 8-byte pointers
    int pool = stride - ar::FIXED;
    if (pool > MAXC * 4 * NCOL) pool = MAXC * 4 * NCOL;
    if (pool < ar::MIN_POOL) { delete m; return fail("rsrx_model_create: not enough shared memory per block for the arena"); }
    m->host.arena_stride = (ar::FIXED + pool + 3) & ~3;
    // tests: RSRX_POOL_LIMIT caps what contacts may allocate in the pool (0 = every contact's rows go to the spill row)
    if (const char* f = getenv("RSRX_POOL_LIMIT")) pool = std::min(pool, std::max(0, atoi(f)));
    m->host.pool_floats = pool;
    m->smem_bytes = WPB * m->host.arena_stride * (int)sizeof(float);
    // tests: RSRX_CONTACT_CAP lowers the number of active contacts the fast kernel accepts before it hands the env to
    // the large-capacity kernel (0 = every env-step with a contact goes there)
    m->host.contact_cap = MAXC;
    if (const char* f = getenv("RSRX_CONTACT_CAP")) m->host.contact_cap = std::min(MAXC, std::max(0, atoi(f)));
    const char* err = nullptr;
    int big_maxc = 0;
    if (rsrx_big_prepare(&m->host, sizeof(DModel), max_smem, &m->big_dev, &m->big_smem, &big_maxc, &err)) {
      std::string msg = err ? err : "rsrx_big_prepare failed";
      delete m;
      return fail(msg);
    }
    m->big_grid = m->num_sms;
    // not when a test knob reshapes the fast kernel (they are about its spill / redo / launch-shape paths), nor with RSRX_MID=0
    const char* mid_off = getenv("RSRX_MID");
    if (!(mid_off && mid_off[0] == '0') && !getenv("RSRX_POOL_LIMIT") && !getenv("RSRX_CONTACT_CAP") && !getenv("RSRX_FORCE_WPB")) {
      if (rsrx_mid_prepare(&m->host, sizeof(DModel), max_smem, &m->mid_dev, &m->mid_arena_bytes, &err)) {
        std::string msg = err ? err : "rsrx_mid_prepare failed";
        if (m->big_dev) cudaFree(m->big_dev);
        delete m;
        return fail(msg);
      }
    }
  }
  cudaError_t e = cudaMalloc(&m->dev, sizeof(DModel));
  if (e == cudaSuccess) e = cudaMemcpy(m->dev, &m->host, sizeof(DModel), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(step_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(step_kernel<kSmallW>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmallW * m->host.arena_stride * (int)sizeof(float));
  if (e == cudaSuccess) e = cudaFuncSetAttribute(reset_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(physics_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, m->smem_bytes);
  if (e != cudaSuccess) {
    std::string msg = std::string("rsrx_model_create: ") + cudaGetErrorString(e);
    if (m->dev) cudaFree(m->dev);
    delete m;
    return fail(msg);
  }
  *out = m;
  return 0;
}

extern "C" void rsrx_model_destroy(rsrx_model* m) {
  if (!m) return;
  if (m->dev) cudaFree(m->dev);
  if (m->big_dev) cudaFree(m->big_dev);
  if (m->mid_dev) cudaFree(m->mid_dev);
  if (m->spill) cudaFree(m->spill);
  if (m->redo) cudaFree(m->redo);
  for (void* p : m->retired) cudaFree(p);
  delete m;
}

extern "C" int rsrx_model_layout(const rsrx_model* m, rsrx_layout* out) {
  if (!m || !out) return fail("rsrx_model_layout: null argument");
  *out = m->host.lay;
  return 0;
}

static PerEnv to_pe(const rsrx_model* m, const rsrx_per_env* p) {
  PerEnv pe = {nullptr, nullptr, nullptr, nullptr, m->spill, m->redo};
  if (p) { pe.geom_friction = p->geom_friction; pe.body_mass = p->body_mass; pe.dof_damping = p->dof_damping; pe.dof_frictionloss = p->dof_frictionloss; }
  return pe;
}
// make sure the spill buffer and the redo list cover N envs (the only allocation the library makes after model
// creation: the first call that sees a larger N, never inside a stream capture)
static int ensure_spill(const rsrx_model* cm, int N) {
  rsrx_model* m = const_cast<rsrx_model*>(cm);
  if (N <= m->spill_envs) return 0;
  float* fresh = nullptr;
  int* redo = nullptr;
  cudaError_t e = cudaMalloc(&fresh, sizeof(float) * (size_t)N * ar::SPILL_STRIDE);
  if (e == cudaSuccess) e = cudaMalloc(&redo, sizeof(int) * ((size_t)N + 2));
  if (e == cudaSuccess) e = cudaMemset(redo, 0, sizeof(int) * ((size_t)N + 2));
  if (e != cudaSuccess) {
    if (fresh) cudaFree(fresh);
    if (redo) cudaFree(redo);
    return fail(std::string("rsrx: cannot allocate the Jacobian spill buffer (call reset/step once for this batch size "
                            "before capturing a CUDA graph): ") + cudaGetErrorString(e));
  }
  if (m->spill) { m->retired.push_back(m->spill); m->retired.push_back(m->redo); }
  m->spill = fresh;
  m->redo = redo;
  m->spill_envs = N;
  return 0;
}
// the large-capacity pass over whatever the fast kernel left on the redo list (normally nothing: ~3 us)
static cudaError_t launch_redo(const rsrx_model* m, rsrx_redo_launch a, const rsrx_per_env* p, const rsrx_state* st, cudaStream_t s) {
  a.geom_friction = p ? p->geom_friction : nullptr; a.body_mass = p ? p->body_mass : nullptr;
  a.dof_damping = p ? p->dof_damping : nullptr; a.dof_frictionloss = p ? p->dof_frictionloss : nullptr;
  a.redo = m->redo;
  if (st) {
    a.data = st->data; a.first_data = st->first_data; a.obs = st->obs; a.first_obs = st->first_obs; a.reward = st->reward;
    a.done = st->done; a.info = st->info; a.metrics = st->metrics; a.status = st->status;
  }
  return rsrx_big_launch(m->big_dev, a, m->big_grid, m->big_smem, s);
}
static rsrx_redo_launch redo_args(int mode) {
  rsrx_redo_launch a;
  memset(&a, 0, sizeof(a));
  a.mode = mode;
  return a;
}
static StatePtrs to_sp(const rsrx_state& s) {
  StatePtrs p;
  p.data = s.data; p.first_data = s.first_data; p.obs = s.obs; p.first_obs = s.first_obs; p.reward = s.reward; p.done = s.done;
  p.info = s.info; p.metrics = s.metrics; p.status = s.status;
  return p;
}
static int check_state(const rsrx_state& s) {
  if (!s.data || !s.first_data || !s.obs || !s.first_obs || !s.reward || !s.done || !s.info || !s.metrics || !s.status)
    return fail("rsrx_state: every buffer must be non-null");
  return 0;
}

extern "C" int rsrx_env_reset(const rsrx_model* m, int N, const float* qpos, const float* qvel, const float* ctrl,
                              const rsrx_per_env* per_env, rsrx_state st, void* stream) {
  if (!m || !qpos || !qvel || !ctrl) return fail("rsrx_env_reset: null argument");
  if (N <= 0) return fail("rsrx_env_reset: N must be positive");
  if (check_state(st)) return 1;
  if (ensure_spill(m, N)) return 1;
  const LaunchCfg lc = launch_cfg(m, N);
  reset_kernel<<<lc.grid, lc.block, lc.smem, (cudaStream_t)stream>>>(m->dev, N, qpos, qvel, ctrl, to_pe(m, per_env), to_sp(st));
  CUDA_OK(cudaGetLastError());
  rsrx_redo_launch ra = redo_args(0);
  ra.qpos = qpos; ra.qvel = qvel; ra.ctrl = ctrl;
  CUDA_OK(launch_redo(m, ra, per_env, &st, (cudaStream_t)stream));
  return 0;
}

// step_kernel (the mid-capacity instantiation for small batches, else the fast one) + the redo pass
static int launch_step(const rsrx_model* m, int N, const float* action, const rsrx_per_env* per_env, const rsrx_state& st,
                       cudaStream_t s) {
  if (ensure_spill(m, N)) return 1;
  if (m->mid_dev && N <= m->num_sms * kMidWarps) {
    rsrx_mid_launch a;
    a.N = N; a.action = action;
    a.geom_friction = per_env ? per_env->geom_friction : nullptr; a.body_mass = per_env ? per_env->body_mass : nullptr;
    a.dof_damping = per_env ? per_env->dof_damping : nullptr; a.dof_frictionloss = per_env ? per_env->dof_frictionloss : nullptr;
    a.redo = m->redo;
    a.data = st.data; a.first_data = st.first_data; a.obs = st.obs; a.first_obs = st.first_obs; a.reward = st.reward;
    a.done = st.done; a.info = st.info; a.metrics = st.metrics; a.status = st.status;
    CUDA_OK(rsrx_mid_launch_step(m->mid_dev, a, m->num_sms, m->mid_arena_bytes, s));
  } else {
    const LaunchCfg lc = launch_cfg(m, N);
    if (lc.block <= 32 * kSmallW) step_kernel<kSmallW><<<lc.grid, lc.block, lc.smem, s>>>(m->dev, N, action, to_pe(m, per_env), to_sp(st));
    else step_kernel<WPB><<<lc.grid, lc.block, lc.smem, s>>>(m->dev, N, action, to_pe(m, per_env), to_sp(st));
    CUDA_OK(cudaGetLastError());
  }
  rsrx_redo_launch ra = redo_args(1);
  ra.action = action;
  CUDA_OK(launch_redo(m, ra, per_env, &st, s));
  return 0;
}

extern "C" int rsrx_env_step(const rsrx_model* m, int N, rsrx_state st, const float* action, const rsrx_per_env* per_env,
                             void* stream) {
  if (!m || !action) return fail("rsrx_env_step: null argument");
  if (N <= 0) return fail("rsrx_env_step: N must be positive");
  if (check_state(st)) return 1;
  return launch_step(m, N, action, per_env, st, (cudaStream_t)stream);
}

extern "C" int rsrx_env_step_host(const rsrx_model* m, int N, rsrx_state st, const float* host_action,
                                  float* action_staging, float* host_obs, float* host_reward, float* host_done,
                                  const rsrx_per_env* per_env, void* stream) {
  if (!m || !host_action || !action_staging) return fail("rsrx_env_step_host: null argument");
  if (N <= 0) return fail("rsrx_env_step_host: N must be positive");
  if (check_state(st)) return 1;
  cudaStream_t s = (cudaStream_t)stream;
  const rsrx_layout& L = m->host.lay;
  CUDA_OK(cudaMemcpyAsync(action_staging, host_action, sizeof(float) * (size_t)N * m->host.nu, cudaMemcpyHostToDevice, s));
  if (launch_step(m, N, action_staging, per_env, st, s)) return 1;
  if (host_obs) CUDA_OK(cudaMemcpyAsync(host_obs, st.obs, sizeof(float) * (size_t)N * L.obs_stride, cudaMemcpyDeviceToHost, s));
  if (host_reward) CUDA_OK(cudaMemcpyAsync(host_reward, st.reward, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, s));
  if (host_done) CUDA_OK(cudaMemcpyAsync(host_done, st.done, sizeof(float) * (size_t)N, cudaMemcpyDeviceToHost, s));
  return 0;
}

extern "C" int rsrx_physics_step(const rsrx_model* m, int N, float* data, int nsteps, const rsrx_per_env* per_env,
                                 int32_t* status, void* stream) {
  if (!m || !data) return fail("rsrx_physics_step: null argument");
  if (N <= 0 || nsteps < 0) return fail("rsrx_physics_step: bad N / nsteps");
  if (ensure_spill(m, N)) return 1;
  const LaunchCfg lc = launch_cfg(m, N);
  physics_kernel<<<lc.grid, lc.block, lc.smem, (cudaStream_t)stream>>>(m->dev, N, data, nsteps, to_pe(m, per_env), status, nullptr);
  CUDA_OK(cudaGetLastError());
  rsrx_redo_launch ra = redo_args(2);
  ra.phys_data = data; ra.nsteps = nsteps; ra.phys_status = status;
  CUDA_OK(launch_redo(m, ra, per_env, nullptr, (cudaStream_t)stream));
  return 0;
}

extern "C" int rsrx_physics_step_debug(const rsrx_model* m, int N, float* data, const rsrx_per_env* per_env, float* dump,
                                       void* stream) {
  if (!m || !data || !dump) return fail("rsrx_physics_step_debug: null argument");
  if (N <= 0) return fail("rsrx_physics_step_debug: bad N");
  if (ensure_spill(m, N)) return 1;
  const LaunchCfg lc = launch_cfg(m, N);
  physics_kernel<<<lc.grid, lc.block, lc.smem, (cudaStream_t)stream>>>(m->dev, N, data, 1, to_pe(m, per_env), nullptr, dump);
  CUDA_OK(cudaGetLastError());
  return 0;
}

// ---- RSR loss -------------------------------------------------------------------
extern "C" int rsrx_kde(const float* grid, int M, int D, const float* data, int Ndata, float bandwidth, float* density_out,
                        void* stream) {
  if (!grid || !data || !density_out) return fail("rsrx_kde: null argument");
  if (M <= 0 || M > loss::MAXM || D <= 0 || Ndata <= 0 || !(bandwidth > 0.f)) return fail("rsrx_kde: bad sizes (M <= 64)");
  return loss::launch(grid, M, D, nullptr, 0, data, Ndata, nullptr, bandwidth, 0.f, 0.f, density_out, nullptr, nullptr,
                      (cudaStream_t)stream)
             ? fail(std::string("rsrx_kde: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_debug_narrowphase(const float* pairs, int n, int plane, float* out, void* stream) {
  if (!pairs || !out || n <= 0) return fail("rsrx_debug_narrowphase: bad arguments");
  narrowphase_kernel<<<(n + 1) / 2, 32, 0, (cudaStream_t)stream>>>(pairs, n, plane, out);
  CUDA_OK(cudaGetLastError());
  return 0;
}

extern "C" int rsrx_tanh_normal_act(const float* logits, const float* noise, int N, int A, float* raw_action, float* action,
                                    float* log_prob, void* stream) {
  if (!logits || !noise || !raw_action || !action) return fail("rsrx_tanh_normal_act: null argument");
  if (N <= 0 || A <= 0) return fail("rsrx_tanh_normal_act: bad sizes");
  return ppo::launch_act(logits, noise, N, A, raw_action, action, log_prob, (cudaStream_t)stream)
             ? fail(std::string("rsrx_tanh_normal_act: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_gather_rows(const float* const* src, float* const* dst, const int32_t* row_floats, int nfields,
                                const int64_t* idx, int nrows, void* stream) {
  if (!src || !dst || !row_floats || !idx) return fail("rsrx_gather_rows: null argument");
  if (nfields <= 0 || nfields > gather::MAXF || nrows <= 0) return fail("rsrx_gather_rows: 1..8 fields, nrows > 0");
  gather::Fields f;
  f.n = nfields;
  for (int k = 0; k < nfields; k++) {
    if (!src[k] || !dst[k] || row_floats[k] <= 0) return fail("rsrx_gather_rows: bad field");
    f.src[k] = src[k]; f.dst[k] = dst[k]; f.width[k] = row_floats[k];
  }
  return gather::launch(f, reinterpret_cast<const long long*>(idx), nrows, (cudaStream_t)stream)
             ? fail(std::string("rsrx_gather_rows: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" size_t rsrx_act_bias_backward_workspace(int rows, int cols) { return mlp::workspace_floats(rows, cols); }

extern "C" int rsrx_act_bias_backward(const float* grad_y, const float* z, int rows, int cols, int activation, float* grad_z,
                                      float* grad_bias, float* workspace, void* stream) {
  if (!grad_y || !grad_bias || !workspace || (activation != 0 && !z)) return fail("rsrx_act_bias_backward: null argument");
  if (rows <= 0 || cols <= 0 || activation < 0 || activation > 2) return fail("rsrx_act_bias_backward: bad sizes / activation");
  return mlp::launch(grad_y, z, rows, cols, activation, grad_z, grad_bias, workspace, (cudaStream_t)stream)
             ? fail(std::string("rsrx_act_bias_backward: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_ppo_head(const float* logits, const float* baseline, const float* bootstrap_value,
                             const float* raw_action, const float* behaviour_log_prob, const float* reward,
                             const float* discount, const float* truncation, const float* noise, int B, int T, int A,
                             float reward_scaling, float discounting, float gae_lambda, float clipping_epsilon,
                             float entropy_cost, int normalize_advantage, float* workspace, float* out,
                             float* grad_logits, float* grad_baseline, void* stream) {
  if (!logits || !baseline || !bootstrap_value || !raw_action || !behaviour_log_prob || !reward || !discount ||
      !truncation || !noise || !workspace || !out || !grad_logits || !grad_baseline)
    return fail("rsrx_ppo_head: null argument");
  if (B <= 0 || T <= 0 || A <= 0 || A > ppo::MAXA) return fail("rsrx_ppo_head: bad sizes (1 <= A <= 16)");
  ppo::Hyper h{reward_scaling, discounting, gae_lambda, clipping_epsilon, entropy_cost, normalize_advantage};
  return ppo::launch(logits, baseline, bootstrap_value, raw_action, behaviour_log_prob, reward, discount, truncation,
                     noise, B, T, A, h, workspace, out, grad_logits, grad_baseline, (cudaStream_t)stream)
             ? fail(std::string("rsrx_ppo_head: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_rsr_loss(const float* grid, int M, int D, const float* reference_data, int Nref, const float* batch,
                             int Nb, const float* reference_density, float bandwidth, float divergence, float loss_scale,
                             float* density_out, float* out, float* grad_batch, void* stream) {
  if (!grid || !batch || !reference_density || !out) return fail("rsrx_rsr_loss: null argument");
  if (Nref > 0 && !reference_data) return fail("rsrx_rsr_loss: reference_data is null");
  if (M <= 0 || M > loss::MAXM || D <= 0 || Nb <= 0 || Nref < 0 || !(bandwidth > 0.f)) return fail("rsrx_rsr_loss: bad sizes (M <= 64)");
  return loss::launch(grid, M, D, reference_data, Nref, batch, Nb, reference_density, bandwidth, divergence, loss_scale,
                      density_out, out, grad_batch, (cudaStream_t)stream)
             ? fail(std::string("rsrx_rsr_loss: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_rsr_policy_term(const float* grid, int M, const float* reference_data, int Nref, const float* reference_density,
                                    float bandwidth, float divergence, float loss_scale, const float* obs, const float* logits,
                                    const float* next_obs, int rows, int O, int A, float* transition, float* grad_transition,
                                    float* out, void* stream) {
  if (!grid || !reference_density || !obs || !logits || !next_obs || !transition || !grad_transition || !out)
    return fail("rsrx_rsr_policy_term: null argument");
  if (Nref > 0 && !reference_data) return fail("rsrx_rsr_policy_term: reference_data is null");
  const int D = 2 * O + A;
  if (M <= 0 || M > loss::MAXM || rows <= 0 || O <= 0 || A <= 0 || D > loss::MAXD || Nref < 0 || !(bandwidth > 0.f))
    return fail("rsrx_rsr_policy_term: bad sizes (M <= 64, 2 O + A <= 256)");
  const int blocks = (int)std::min<size_t>(((size_t)rows * D + 255) / 256, 4 * 148);
  CUDA_OK(pdl::launch(loss::pack_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, obs, logits, next_obs, rows, O, A, transition));
  return loss::launch(grid, M, D, reference_data, Nref, transition, rows, reference_density, bandwidth, divergence, loss_scale, nullptr,
                      out, grad_transition, (cudaStream_t)stream)
             ? fail(std::string("rsrx_rsr_policy_term: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}

extern "C" int rsrx_rsr_logit_grad(const float* transition, const float* grad_transition, const float* grad_logits_in, int rows, int O,
                                   int A, float* grad_logits_out, void* stream) {
  if (!transition || !grad_transition || !grad_logits_out || rows <= 0 || O <= 0 || A <= 0)
    return fail("rsrx_rsr_logit_grad: bad arguments");
  const int blocks = (int)std::min<size_t>(((size_t)rows * 2 * A + 255) / 256, 4 * 148);
  CUDA_OK(pdl::launch(loss::unpack_kernel, dim3(blocks), dim3(256), 0, (cudaStream_t)stream, transition, grad_transition, grad_logits_in,
                      rows, O, A, grad_logits_out));
  return 0;
}

// ---- tensor-core linear layers (csrc/rsrx_gemm.cuh) -----------------------------------------------------------------
static unsigned long long* gemm_dbg() {  // RSRX_GEMM_STAMPS = device address of 8 uint64 (profiling aid)
  const char* d = getenv("RSRX_GEMM_STAMPS");
  return d ? reinterpret_cast<unsigned long long*>(strtoull(d, nullptr, 10)) : nullptr;
}
static int gemm_check(const char* what, const void* a, const void* b, int lda, int ldb) {
  if (((uintptr_t)a & 15) || ((uintptr_t)b & 15) || (lda & 3) || (ldb & 3))
    return fail(std::string(what) + ": operands must be 16-byte aligned with leading dimensions that are multiples of 4");
  return 0;
}

extern "C" int rsrx_linear_forward(const float* x, int ldx, const float* w, int ldw, const float* bias, int M, int N, int K,
                                   int activation, float* z, float* y, int ldy, float* yT, int ldt, void* stream) {
  if (!x || !w || !y) return fail("rsrx_linear_forward: null argument");
  if (M <= 0 || N <= 0 || K <= 0 || (K & 3) || activation < 0 || activation > 2) return fail("rsrx_linear_forward: bad sizes (K % 4 == 0)");
  if (gemm_check("rsrx_linear_forward", x, w, ldx, ldw)) return 1;
  gemm::Params p{};
  p.A = x; p.a_row = ldx; p.a_col = 1; p.B = w; p.b_row = ldw; p.b_col = 1; p.M = M; p.N = N; p.K = K; p.k_split = K;
  p.epilogue = gemm::EPI_BIAS_ACT; p.act = activation; p.bias = bias; p.D = y; p.Z = z; p.ldd = ldy; p.DT = yT; p.ldt = ldt;
  if (((uintptr_t)y & 15) || ((uintptr_t)z & 15) || (ldy & 3)) return fail("rsrx_linear_forward: outputs must be 16-byte aligned, ldy % 4 == 0");
  if (yT && ldt < M) return fail("rsrx_linear_forward: ldt < M");
  p.dbg_t = gemm_dbg();
  CUDA_OK(gemm::launch(p, 0, (cudaStream_t)stream));
  return 0;
}

extern "C" int rsrx_linear_dgrad(const float* dz, int lddz, const float* w, int ldw, const float* wT, int ldwt, const float* zprev,
                                 int M, int Nin, int Nout, int activation, float* dzprev, int ld, float* colsum_partials,
                                 float* dzprevT, int ldt, void* stream) {
  if (!dz || (!w && !wT) || !dzprev) return fail("rsrx_linear_dgrad: null argument");
  if (M <= 0 || Nin <= 0 || Nout <= 0 || (Nout & 3) || (Nin & 3) || activation < 0 || activation > 2 || (activation && !zprev))
    return fail("rsrx_linear_dgrad: bad sizes");
  if (gemm_check("rsrx_linear_dgrad", dz, wT ? wT : w, lddz, wT ? ldwt : ldw)) return 1;
  gemm::Params p{};
  p.A = dz; p.a_row = lddz; p.a_col = 1;          // [M][Nout], contraction over the layer's outputs
  if (wT) { p.B = wT; p.b_row = ldwt; p.b_col = 1; }  // wT[k'][n']: contiguous along the contraction -> TMA
  else { p.B = w; p.b_row = 1; p.b_col = ldw; }       // B[k'][n'] = w[n'][k']: transposed while staged
  p.M = M; p.N = Nin; p.K = Nout; p.k_split = Nout;
  p.epilogue = gemm::EPI_DGRAD; p.act = activation; p.zprev = zprev; p.D = dzprev; p.colsum = colsum_partials; p.ldd = ld;
  p.DT = dzprevT; p.ldt = ldt;
  if (((uintptr_t)dzprev & 15) || ((uintptr_t)zprev & 15) || (ld & 3)) return fail("rsrx_linear_dgrad: outputs must be 16-byte aligned, ld % 4 == 0");
  if (dzprevT && ldt < M) return fail("rsrx_linear_dgrad: ldt < M");
  p.dbg_t = gemm_dbg();
  CUDA_OK(gemm::launch(p, wT ? 0 : 1, (cudaStream_t)stream));
  return 0;
}

extern "C" int rsrx_linear_wgrad(const float* dz, int lddz, const float* x, int ldx, int transposed_inputs, int rows, int Nout,
                                 int Nin, int rows_per_split, float* partials, int ldp, void* stream) {
  if (!dz || !x || !partials) return fail("rsrx_linear_wgrad: null argument");
  if (rows <= 0 || Nout <= 0 || Nin <= 0 || (Nout & 3) || (Nin & 3) || rows_per_split <= 0 || (rows_per_split % gemm::BK))
    return fail("rsrx_linear_wgrad: bad sizes (rows_per_split % 32 == 0)");
  if (gemm_check("rsrx_linear_wgrad", dz, x, lddz, ldx)) return 1;
  gemm::Params p{};
  if (transposed_inputs) {  // dz = dZ^T [Nout][lddz >= rows], x = X^T [Nin][ldx >= rows]: contiguous along the contraction -> TMA
    p.A = dz; p.a_row = lddz; p.a_col = 1;
    p.B = x; p.b_row = ldx; p.b_col = 1;
  } else {
    p.A = dz; p.a_row = 1; p.a_col = lddz;        // A[n'][m] = dz[m][n']
    p.B = x; p.b_row = 1; p.b_col = ldx;          // B[k'][m] = x[m][k']
  }
  p.M = Nout; p.N = Nin; p.K = rows; p.k_split = rows_per_split;
  p.epilogue = gemm::EPI_PARTIAL; p.D = partials; p.ldd = ldp;
  if (((uintptr_t)partials & 15) || (ldp & 3)) return fail("rsrx_linear_wgrad: partials must be 16-byte aligned, ldp % 4 == 0");
  p.dbg_t = gemm_dbg();
  CUDA_OK(gemm::launch(p, transposed_inputs ? 0 : 2, (cudaStream_t)stream));
  return 0;
}

extern "C" int rsrx_reduce_partials(const float* const* in, float* const* out, const int32_t* n, const int32_t* S,
                                    const int64_t* stride, int nseg, void* stream) {
  if (!in || !out || !n || !S || !stride || nseg <= 0 || nseg > gemm::MAXSEG) return fail("rsrx_reduce_partials: 1..24 segments");
  gemm::ReduceArgs a;
  a.nseg = nseg;
  int nmax = 0;
  for (int k = 0; k < nseg; k++) {
    if (!in[k] || !out[k] || n[k] <= 0 || S[k] <= 0) return fail("rsrx_reduce_partials: bad segment");
    a.seg[k] = {in[k], out[k], n[k], S[k], (long long)stride[k]};
    nmax = std::max(nmax, n[k]);
  }
  const dim3 grid(std::min((nmax + 255) / 256, 256), nseg);  // one output per thread on the 256 x 256 weight gradients
  CUDA_OK(pdl::launch(gemm::reduce_partials_kernel, grid, dim3(256), 0, (cudaStream_t)stream, a));
  return 0;
}

extern "C" int rsrx_value_head_backward(const float* g, const float* w, const float* z, const float* h, int M, int n, int ld,
                                        int activation, float* dz, float* colsum_partials, float* dw_partials,
                                        float* db_partials, float* dzT, int ldt, void* stream) {
  if (!g || !w || !h || !dz || !colsum_partials || !dw_partials || !db_partials || (activation && !z))
    return fail("rsrx_value_head_backward: null argument");
  if (M <= 0 || n <= 0 || ld < n || activation < 0 || activation > 2 || (dzT && ldt < M)) return fail("rsrx_value_head_backward: bad sizes");
  if ((n & 3) || (ld & 3)) return fail("rsrx_value_head_backward: n and ld must be multiples of 4");
  CUDA_OK(pdl::launch(gemm::head_backward_kernel, dim3((M + 127) / 128, (n + 63) / 64), dim3(256), 0, (cudaStream_t)stream, g, w, z, h, M, n,
                      ld, activation, dz, colsum_partials, dw_partials, db_partials, dzT, ldt));
  return 0;
}

extern "C" int rsrx_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                              const int32_t* sizes, float* const* params_t, const int32_t* cols, int ntensors, float lr,
                              float beta1, float beta2, float eps, float grad_scale, uint64_t* step_ticket, void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || !sizes || !step_ticket) return fail("rsrx_adam_step: null argument");
  if (ntensors <= 0 || ntensors > gemm::ADAM_MAXSEG) return fail("rsrx_adam_step: 1..32 tensors");
  gemm::AdamArgs a;
  a.nseg = ntensors; a.lr = lr; a.beta1 = beta1; a.beta2 = beta2; a.eps = eps; a.grad_scale = grad_scale;
  a.counters = reinterpret_cast<unsigned int*>(step_ticket);  // {steps taken, blocks done}
  for (int k = 0; k < ntensors; k++) {
    if (!params[k] || !grads[k] || !exp_avg[k] || !exp_avg_sq[k] || sizes[k] <= 0) return fail("rsrx_adam_step: bad tensor");
    float* pt = params_t ? params_t[k] : nullptr;
    const int c = (pt && cols) ? cols[k] : 1;
    if (pt && (c <= 0 || sizes[k] % c)) return fail("rsrx_adam_step: bad transposed-copy shape");
    a.seg[k] = {params[k], grads[k], exp_avg[k], exp_avg_sq[k], sizes[k], pt, c};
  }
  const dim3 grid(gemm::ADAM_BLOCKS_X, ntensors);
  CUDA_OK(pdl::launch(gemm::adam_kernel, grid, dim3(256), 0, (cudaStream_t)stream, a));
  return 0;
}

// ---- warp-per-row small MLP (csrc/rsrx_mlp.cuh): the trainers' policy network --------------------------------------
static int small_net(const float* const* weights, const float* const* biases, const int32_t* widths, int nlayers, int activation,
                     smallmlp::Net& net, int* total) {
  if (!weights || !biases || !widths) return fail("rsrx_small_mlp: null argument");
  if (nlayers <= 0 || nlayers > smallmlp::MAXL || activation < 0 || activation > 2) return fail("rsrx_small_mlp: 1..8 layers, activation 0..2");
  int t = 0;
  for (int l = 0; l <= nlayers; l++)
    if (widths[l] <= 0 || widths[l] > smallmlp::WD) return fail("rsrx_small_mlp: every width must be 1..32");
  for (int l = 0; l < nlayers; l++) {
    if (!weights[l] || !biases[l]) return fail("rsrx_small_mlp: null parameter");
    net.W[l] = weights[l]; net.b[l] = biases[l];
    t += widths[l] * widths[l + 1] + widths[l + 1];
  }
  for (int l = 0; l <= nlayers; l++) net.width[l] = widths[l];
  net.nl = nlayers; net.act = activation;
  if (total) *total = t;
  return 0;
}
// CTAs of the warp-per-row kernels: a row is a ~1000-instruction dependent chain on its warp, so the rows are spread as
// widely as the chip allows (one CTA per SM at 2560 rows: ~2 rows per warp; 40 CTAs of 8 rows per warp measured 94 vs 51 us)
static int small_grid(int rows, int rows_per_warp) {
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return std::max(1, std::min(sms, (rows + smallmlp::WARPS * rows_per_warp - 1) / (smallmlp::WARPS * rows_per_warp)));
}
extern "C" int rsrx_small_mlp_backward_ctas(int rows) { return small_grid(rows, 2); }

extern "C" int rsrx_small_mlp_forward(const float* const* weights, const float* const* biases, const int32_t* widths, int nlayers,
                                      int activation, const float* x, int ldx, int rows, float* zs, float* out, int ldo,
                                      const float* norm_mean, const float* norm_std, void* stream) {
  smallmlp::Net net;
  if (small_net(weights, biases, widths, nlayers, activation, net, nullptr)) return 1;
  if (!x || !out || (nlayers > 1 && !zs) || rows <= 0) return fail("rsrx_small_mlp_forward: bad arguments");
  if ((norm_mean == nullptr) != (norm_std == nullptr)) return fail("rsrx_small_mlp_forward: norm_mean and norm_std go together");
  const size_t smem = smallmlp::fwd_smem(nlayers);
  static bool set = false;
  if (!set) { CUDA_OK(cudaFuncSetAttribute(smallmlp::forward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smallmlp::fwd_smem(smallmlp::MAXL))); set = true; }
  CUDA_OK(pdl::launch(smallmlp::forward_kernel, dim3(small_grid(rows, 2)), dim3(32 * smallmlp::WARPS), smem, (cudaStream_t)stream, net, x, ldx,
                      rows, zs, out, ldo, norm_mean, norm_std));
  return 0;
}

extern "C" int rsrx_small_mlp_backward(const float* const* weights, const float* const* biases, const int32_t* widths, int nlayers,
                                       int activation, const float* x, int ldx, int rows, const float* zs, const float* grad_out,
                                       int ldg, float* partials, void* stream) {
  smallmlp::Net net;
  int total = 0;
  if (small_net(weights, biases, widths, nlayers, activation, net, &total)) return 1;
  if (!x || !grad_out || !partials || (nlayers > 1 && !zs) || rows <= 0) return fail("rsrx_small_mlp_backward: bad arguments");
  const size_t smem = smallmlp::bwd_smem(nlayers);
  if (smem > 227 * 1024) return fail("rsrx_small_mlp_backward: too many layers for the shared-memory accumulators");
  static bool set = false;
  if (!set) { CUDA_OK(cudaFuncSetAttribute(smallmlp::backward_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024)); set = true; }
  CUDA_OK(pdl::launch(smallmlp::backward_kernel, dim3(small_grid(rows, 2)), dim3(32 * smallmlp::WARPS), smem, (cudaStream_t)stream, net, x,
                      ldx, rows, zs, grad_out, ldg, partials, total));
  return 0;
}

extern "C" int rsrx_ppo_prep(const float* obs, const float* next_obs, const float* mean, const float* std, int mb, int T, int O,
                             float* obs_n, float* x_pad, int ldp, float* xT, int ldt, void* stream) {
  if (!obs || !next_obs || !mean || !std || !obs_n || !x_pad || !xT) return fail("rsrx_ppo_prep: null argument");
  if (mb <= 0 || T <= 0 || O <= 0 || ldp < O || ldt < mb * T + mb) return fail("rsrx_ppo_prep: bad sizes");
  return ppo::launch_prep(obs, next_obs, mean, std, mb, T, O, obs_n, x_pad, ldp, xT, ldt, (cudaStream_t)stream)
             ? fail(std::string("rsrx_ppo_prep: ") + cudaGetErrorString(cudaGetLastError()))
             : 0;
}
