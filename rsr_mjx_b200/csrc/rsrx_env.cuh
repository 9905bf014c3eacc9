// rsrx_env.cuh — fused environment kernels: one launch = one brax
// AutoResetWrapper(EpisodeWrapper(VmapWrapper(env))).step for N envs.
//   reference env code: test/airbot.py:165-268, cube_env.py:145-229,
//   T_shape_env.py:139-234; wrappers: mujoco_playground/_src/wrapper.py:117-138.
#pragma once
#include "rsrx_physics.cuh"

namespace RSRX_NS {

struct PerEnv {
  const float* geom_friction;
  const float* body_mass;
  const float* dof_damping;
  const float* dof_frictionloss;
  float* spill;  // [slots][ar::SPILL_STRIDE] library-owned overflow of the Jacobian-row pool
  // redo list (library-owned): [0] = number of envs this launch could not finish within contact_cap, [1] = CTA ticket
  // of the redo kernel, [2 + i] = env ids.  nullptr: no redo, the env only gets RSRX_STATUS_CONTACT_OVERFLOW.
  int* redo;
};

// load the dynamic state of env `e` into the arena; per-env model arrays stay in global memory and
// are reached through pointers (pre-offset to this env) kept at ar::PTRS
__device__ __noinline__ void load_env(const DModel* __restrict__ dm, float* sm, int lane, int e, int spill_slot,
                                      const float* __restrict__ row, const PerEnv& pe) {
  const rsrx_layout& L = dm->lay;
#pragma unroll 1
  for (int i = lane; i < dm->nq; i += 32) sm[ar::QPOS + i] = row[L.qpos + i];
#pragma unroll 1
  for (int i = lane; i < dm->nv; i += 32) {
    sm[ar::QVEL + i] = row[L.qvel + i];
    sm[ar::WARM + i] = row[L.qacc_warmstart + i];
    sm[ar::DAMP + i] = pe.dof_damping ? pe.dof_damping[(size_t)e * dm->nv + i] : dm->dof_damping[i];
  }
#pragma unroll 1
  for (int i = lane; i < dm->nu; i += 32) sm[ar::CTRL + i] = row[L.ctrl + i];
  if (lane == 0) {
    const float** ptrs = reinterpret_cast<const float**>(sm + ar::PTRS);
    ptrs[0] = pe.geom_friction ? pe.geom_friction + (size_t)e * dm->ngeom * 3 : nullptr;
    ptrs[1] = pe.body_mass ? pe.body_mass + (size_t)e * dm->nbody : nullptr;
    ptrs[2] = pe.dof_frictionloss ? pe.dof_frictionloss + (size_t)e * dm->nv : nullptr;
    ptrs[3] = pe.spill + (size_t)spill_slot * ar::SPILL_STRIDE;
    reinterpret_cast<int*>(sm + ar::PTRS)[ar::FLAGS] = 1;
  }
  RSRX_SYNC();
}

// write the pipeline_state row (qpos/qvel/ctrl/warmstart/time + lagged kinematics)
__device__ __noinline__ void store_env(const DModel* __restrict__ dm, const float* sm, int lane, float* __restrict__ row, float time) {
  const rsrx_layout& L = dm->lay;
#pragma unroll 1
  for (int i = lane; i < dm->nq; i += 32) row[L.qpos + i] = sm[ar::QPOS + i];
#pragma unroll 1
  for (int i = lane; i < dm->nv; i += 32) { row[L.qvel + i] = sm[ar::QVEL + i]; row[L.qacc_warmstart + i] = sm[ar::WARM + i]; }
#pragma unroll 1
  for (int i = lane; i < dm->nu; i += 32) row[L.ctrl + i] = sm[ar::CTRL + i];
  if (lane == 0) row[L.time] = time;
#pragma unroll 1
  for (int i = lane; i < dm->nbody * 3; i += 32) row[L.xpos + i] = sm[ar::XPOS + i];
#pragma unroll 1
  for (int i = lane; i < dm->nbody * 4; i += 32) row[L.xquat + i] = sm[ar::XQUAT + i];
#pragma unroll 1
  for (int i = lane; i < dm->nsite * 3; i += 32) row[L.site_xpos + i] = sm[ar::SXPOS + i];
#pragma unroll 1
  for (int g = lane; g < dm->ngeom; g += 32) {
    float gp[3];
    geom_pose(dm, sm, g, gp, nullptr);
    row[L.geom_xpos + g * 3] = gp[0]; row[L.geom_xpos + g * 3 + 1] = gp[1]; row[L.geom_xpos + g * 3 + 2] = gp[2];
  }
}

// _get_obs into the arena's OBSBUF (lane 0); info = this env's info row
__device__ void get_obs(const DModel* __restrict__ dm, float* sm, const float* info) {
  float* obs = sm + ar::OBSBUF;
  int n = 0;
  for (int i = 0; i < 6; i++) obs[n++] = sm[ar::QPOS + dm->joint_qadr[i]];
  const float* site = sm + ar::SXPOS + dm->site_endpoint * 3;
  if (dm->env_kind == RSRX_ENV_T) {
    float gb[3], gv[3];
    geom_pose(dm, sm, dm->geom_base, gb, nullptr);
    geom_pose(dm, sm, dm->geom_vertical, gv, nullptr);
    obs[n++] = site[2];
    for (int i = 0; i < 3; i++) obs[n++] = info[RSRX_INFO_TARGET + i] - gb[i];
    for (int i = 0; i < 3; i++) obs[n++] = info[RSRX_INFO_TARGET2 + i] - gv[i];
    obs[n++] = info[RSRX_INFO_XITA];
    for (int i = 0; i < 2; i++) obs[n++] = info[RSRX_INFO_NEWPOS + i] - site[i];
  } else {
    const float* cube = sm + ar::XPOS + dm->cube_body * 3;
    for (int i = 0; i < 3; i++) obs[n++] = site[i];
    for (int i = 0; i < 3; i++) obs[n++] = info[RSRX_INFO_TARGET + i];
    for (int i = 0; i < 3; i++) obs[n++] = cube[i];
    for (int i = 0; i < 2; i++) obs[n++] = info[RSRX_INFO_NEWPOS + i];
    for (int i = 0; i < 3; i++) obs[n++] = info[RSRX_INFO_TARGET + i] - cube[i];
    for (int i = 0; i < 3; i++) obs[n++] = cube[i] - site[i];
  }
  for (; n < OBS_STRIDE; n++) obs[n] = 0.f;
}


// Warps per CTA: one env per warp, warps are independent (only __syncwarp).  Single-warp CTAs all land on
// the same SM sub-partition (warp id within the CTA selects the scheduler), leaving 3 of the 4 schedulers
// idle — so a CTA carries WPB envs; with the per-substep phase barrier (rsrx_physics.cuh) ONE CTA of
// WPB = 19 warps per SM (what the 12 KB arena allows) keeps every resident warp in the same code region
// (measured at 8192 envs: 1 x 14 warps 4.9 ms, 2 x 7 warps 5.6 ms, 8 warps 6.3 ms, no barrier 9.5 ms).
constexpr int WPB = RSRX_WPB;

struct StatePtrs {
  float *data, *first_data, *obs, *first_obs, *reward, *done, *info, *metrics;
  int* status;
};

// append env e to the redo list (lane 0 of its warp)
__device__ __forceinline__ void redo_push(const PerEnv& pe, int e) {
  if (pe.redo) pe.redo[2 + atomicAdd(pe.redo, 1)] = e;
}

// ------------------------------------------------------------------ env bodies
// One env on one warp.  Each body returns false — having committed NOTHING to the env's state — when a substep saw
// more active contacts than dm->contact_cap: the caller then queues the env for the large-capacity instantiation
// (rsrx_redo.cu), which runs the same body from the same inputs.  status_init is OR-ed into the env's status word.

// reset: pipeline_init (make_data + mjx.forward with ctrl = 0, warmstart = 0), then data.replace(ctrl), info /
// metrics / obs, wrappers' reset bookkeeping.
__device__ __forceinline__ bool env_reset_body(const DModel* __restrict__ dm, float* sm, int lane, int e, int spill_slot,
                                               const float* __restrict__ qpos, const float* __restrict__ qvel,
                                               const float* __restrict__ ctrl, const PerEnv& pe, const StatePtrs& st,
                                               int status_init) {
  const rsrx_layout& L = dm->lay;
  float* row = st.data + (size_t)e * L.data_stride;
  // stage a row: qpos, qvel, ctrl = 0, warm = 0
#pragma unroll 1
  for (int i = lane; i < L.data_stride; i += 32) row[i] = 0.f;
  RSRX_SYNC();
#pragma unroll 1
  for (int i = lane; i < dm->nq; i += 32) row[L.qpos + i] = qpos[(size_t)e * dm->nq + i];
#pragma unroll 1
  for (int i = lane; i < dm->nv; i += 32) row[L.qvel + i] = qvel[(size_t)e * dm->nv + i];
  RSRX_SYNC();
  load_env(dm, sm, lane, e, spill_slot, row, pe);
  SolverDims sd;
  int status = status_init;
  forward<false>(dm, sm, lane, &sd, &status);
  status = __reduce_or_sync(0xffffffffu, status);
  if (status & RSRX_STATUS_CONTACT_OVERFLOW) return false;
#pragma unroll 1
  for (int i = lane; i < dm->nu; i += 32) sm[ar::CTRL + i] = ctrl[(size_t)e * dm->nu + i];
  RSRX_SYNC();
  store_env(dm, sm, lane, row, 0.f);
  float* info = st.info + (size_t)e * RSRX_INFO_STRIDE;
  if (lane == 0) {
    for (int i = 0; i < RSRX_INFO_STRIDE; i++) info[i] = 0.f;
    const float* site = sm + ar::SXPOS + dm->site_endpoint * 3;
    for (int i = 0; i < 3; i++) { info[RSRX_INFO_SITE + i] = site[i]; info[RSRX_INFO_OBJ + i] = sm[ar::XPOS + dm->cube_body * 3 + i]; }
    if (dm->env_kind == RSRX_ENV_T) {
      info[RSRX_INFO_NEWPOS] = 0.24739072f; info[RSRX_INFO_NEWPOS + 1] = -0.00496255f;
      float gb[3], gv[3];
      geom_pose(dm, sm, dm->geom_target_base, gb, nullptr);
      geom_pose(dm, sm, dm->geom_target_vertical, gv, nullptr);
      for (int i = 0; i < 3; i++) { info[RSRX_INFO_TARGET + i] = gb[i]; info[RSRX_INFO_TARGET2 + i] = gv[i]; }
      info[RSRX_INFO_TARGET_W] = sm[ar::XQUAT + dm->target_body * 4] * 10.f;
      info[RSRX_INFO_XITA] = 0.2876f;
    } else {
      info[RSRX_INFO_NEWPOS] = 0.37342f; info[RSRX_INFO_NEWPOS + 1] = -0.07989f;
      for (int i = 0; i < 3; i++) info[RSRX_INFO_TARGET + i] = sm[ar::XPOS + dm->target_body * 3 + i];
    }
    get_obs(dm, sm, info);
    st.reward[e] = 0.f;
    st.done[e] = 0.f;
    for (int i = 0; i < METRICS_STRIDE; i++) st.metrics[(size_t)e * METRICS_STRIDE + i] = 0.f;
    st.status[e] = status;
  }
  RSRX_SYNC();
#pragma unroll 1
  for (int i = lane; i < OBS_STRIDE; i += 32) {
    st.obs[(size_t)e * OBS_STRIDE + i] = sm[ar::OBSBUF + i];
    st.first_obs[(size_t)e * OBS_STRIDE + i] = sm[ar::OBSBUF + i];
  }
  __threadfence_block();
  RSRX_SYNC();
  float* frow = st.first_data + (size_t)e * L.data_stride;
#pragma unroll 1
  for (int i = lane; i < L.data_stride; i += 32) frow[i] = row[i];
  return true;
}

// step: AutoReset pre, action shaping, n_frames x mjx.step, reward / done / info / obs, Episode + AutoReset post
template <bool SYNC>
__device__ __forceinline__ bool env_step_body(const DModel* __restrict__ dm, float* sm, int lane, int e, int spill_slot,
                                              const float* __restrict__ action, const PerEnv& pe, const StatePtrs& st,
                                              int status_init) {
  const rsrx_layout& L = dm->lay;
  const int kind = dm->env_kind;
  float* row = st.data + (size_t)e * L.data_stride;
  float* info = st.info + (size_t)e * RSRX_INFO_STRIDE;
  load_env(dm, sm, lane, e, spill_slot, row, pe);
  float time = row[L.time];
  // ---- AutoReset pre + action shaping (lane 0; reads the *lagged* kinematics of the row).  The two info words this
  // part changes (steps, last_action) stay in registers until the step is known to fit the contact capacity.
  float steps0 = 0.f, last_action = 0.f;
  if (lane == 0) {
    steps0 = info[RSRX_INFO_STEPS];
    if (dm->episode_length > 0 && st.done[e] != 0.f) steps0 = 0.f;
    float act[NU];
    const int nu = dm->nu;
    for (int i = 0; i < nu; i++) act[i] = sm[ar::CTRL + i] + dm->action_scale[i] * action[(size_t)e * nu + i];
    act[3] = -(1.57f + sm[ar::QPOS + dm->joint_qadr[1]] + sm[ar::QPOS + dm->joint_qadr[2]]);
    if (kind == RSRX_ENV_T) {
      const float px = row[L.site_xpos + dm->site_endpoint * 3], py = row[L.site_xpos + dm->site_endpoint * 3 + 1];
      const float dx = row[L.site_xpos + dm->site_tail * 3] - px, dy = row[L.site_xpos + dm->site_tail * 3 + 1] - py;
      act[4] = -atan2f(dy, dx + 0.00001f) + act[0] + 1.5708f;
    } else {
      const float* cube = row + L.xpos + dm->cube_body * 3;
      const float dx = info[RSRX_INFO_TARGET] - cube[0], dy = info[RSRX_INFO_TARGET + 1] - cube[1];
      float a4 = -atan2f(dy, dx + 0.00001f) + act[0] + 1.5708f;
      if (kind == RSRX_ENV_SF) {
        float df[3] = {info[RSRX_INFO_TARGET] - cube[0], info[RSRX_INFO_TARGET + 1] - cube[1], info[RSRX_INFO_TARGET + 2] - cube[2]};
        if (sqrtf(dot3(df, df)) < 0.03f) a4 = info[RSRX_INFO_LAST_ACTION];
        last_action = a4;
      }
      act[4] = a4;
    }
    for (int i = 0; i < nu; i++) sm[ar::CTRL + i] = clipf(act[i], dm->act_ctrlrange[i][0], dm->act_ctrlrange[i][1]);
  }
  RSRX_SYNC();
  // ---- pipeline_step: n_frames x mjx.step
  int status = status_init;
  SolverDims sd;
  for (int f = 0; f < dm->n_frames; ++f) {
    forward<SYNC>(dm, sm, lane, &sd, &status);
    implicit_advance(dm, sm, lane);
    time += dm->timestep;
  }
  status = __reduce_or_sync(0xffffffffu, status);
  if (status & RSRX_STATUS_CONTACT_OVERFLOW) return false;  // nothing written yet: the env-step is redone from the same inputs
  // ---- post-physics (lane 0)
  float done = 0.f;
  if (lane == 0) {
    float reward;
    const float* site = sm + ar::SXPOS + dm->site_endpoint * 3;
    const float W = dm->siet_to_box_reward_weight;
    float* met = st.metrics + (size_t)e * METRICS_STRIDE;
    if (kind == RSRX_ENV_T) {
      float db[3], dv[3], box[3], tgt[3], gb[3], gv[3];
      geom_pose(dm, sm, dm->geom_base, gb, nullptr);
      geom_pose(dm, sm, dm->geom_vertical, gv, nullptr);
      for (int i = 0; i < 3; i++) {
        db[i] = info[RSRX_INFO_TARGET + i] - gb[i];
        dv[i] = info[RSRX_INFO_TARGET2 + i] - gv[i];
        box[i] = gv[i] - gb[i];
        tgt[i] = info[RSRX_INFO_TARGET2 + i] - info[RSRX_INFO_TARGET + i];
      }
      float disb = sqrtf(dot3(db, db)), disv = sqrtf(dot3(dv, dv));
      if (disb < 0.005f) disb = 0.f;
      if (disv < 0.005f) disv = 0.f;
      const float prb = 1.f / (1.f + 10.f * disb), prv = 1.f / (1.f + 10.f * disv);
      const float cosx = dot3(box, tgt) / (sqrtf(dot3(box, box)) * sqrtf(dot3(tgt, tgt)));
      const float xita = acosf(clipf(cosx, -1.f, 1.f));
      info[RSRX_INFO_XITA] = xita;
      const float pwr = 1.f / (1.f + 6.f * xita);
      const float push = (0.1515f * prb + 0.1515f * prv + 0.66f * pwr) * dm->push_reward_weight;
      const float tail0 = sm[ar::SXPOS + dm->site_tail * 3], tail1 = sm[ar::SXPOS + dm->site_tail * 3 + 1];
      const float old0 = info[RSRX_INFO_NEWPOS], old1 = info[RSRX_INFO_NEWPOS + 1];
      float site_z_reward = site[2] < 0.83f ? 1.f : 0.f;
      site_z_reward += 4.f / (1.f + 3.f * fabsf(site[2] - 0.805f));
      const float dx = sm[ar::SXPOS + dm->site_target_tail * 3] - tail0, dy = sm[ar::SXPOS + dm->site_target_tail * 3 + 1] - tail1;
      const float ang = atan2f(dy, dx + 0.00001f);
      const float distance = sqrtf(dx * dx + dy * dy) + 0.025f;
      const float y_ = distance * sinf(ang), x_ = distance * cosf(ang);
      info[RSRX_INFO_NEWPOS] = dx - x_ + tail0;
      info[RSRX_INFO_NEWPOS + 1] = dy - y_ + tail1;
      const float e0 = site[0] - old0, e1 = site[1] - old1;
      float sdis = sqrtf(e0 * e0 + e1 * e1);
      sdis = sdis < 0.02f ? 0.f : sdis - 0.02f;
      const float s2c = (1.f - tanhf(5.f * sdis)) * W;
      const float health = dm->healthy_reward * fabsf((site[2] < dm->endpoint_min_z_pos ? 1.f : 0.f) - 1.f);
      reward = push + s2c + health + site_z_reward;
      done = sm[ar::XPOS + dm->cube_body * 3 + 2] < 0.6f ? 1.f : 0.f;
      met[0] = push; met[1] = s2c; met[2] = health; met[4] = site_z_reward;
    } else {
      const float* cube = sm + ar::XPOS + dm->cube_body * 3;
      float df[3] = {info[RSRX_INFO_TARGET] - cube[0], info[RSRX_INFO_TARGET + 1] - cube[1], info[RSRX_INFO_TARGET + 2] - cube[2]};
      float dis = sqrtf(dot3(df, df));
      const float thr = kind == RSRX_ENV_SF ? 0.003f : 0.005f;
      if (dis < thr) dis = 0.f;
      const float push = (1.f / (1.f + 3.f * dis)) * dm->push_reward_weight;
      const float task_complete = (kind == RSRX_ENV_SF && dis < 0.003f) ? 5.f : 0.f;
      const float old0 = info[RSRX_INFO_NEWPOS], old1 = info[RSRX_INFO_NEWPOS + 1];
      const float site_z_reward = site[2] < 0.82f ? 1.f : 0.f;
      const float dx = info[RSRX_INFO_TARGET] - cube[0], dy = info[RSRX_INFO_TARGET + 1] - cube[1];
      const float ang = atan2f(dy, dx + 0.00001f);
      const float distance = sqrtf(dx * dx + dy * dy) + 0.04f;
      const float y_ = distance * sinf(ang), x_ = distance * cosf(ang);
      info[RSRX_INFO_NEWPOS] = dx - x_ + cube[0];
      info[RSRX_INFO_NEWPOS + 1] = dy - y_ + cube[1];
      const float e0 = site[0] - old0, e1 = site[1] - old1;
      float sdis = sqrtf(e0 * e0 + e1 * e1);
      sdis = sdis < 0.042f ? 0.f : sdis - 0.042f;
      float s2c = (1.f - tanhf(5.f * sdis)) * W;
      if (dis < 0.005f) s2c = W;
      if (kind == RSRX_ENV_SF) {
        float dn = 0.f;
        if (site[2] < dm->endpoint_min_z_pos) dn = 1.f;
        if (site[0] > 1.0f) dn = 1.f;
        if (site[0] < -0.6f) dn = 1.f;
        if (site[1] > 0.3f) dn = 1.f;
        if (site[1] < -0.3f) dn = 1.f;
        if (cube[2] < 0.6f) dn = 1.f;
        const float health = dm->healthy_reward * fabsf(dn - 1.f);
        reward = push + s2c + health + task_complete + site_z_reward;
        done = dis < 0.003f ? 1.f : 0.f;
        met[0] = push; met[1] = 0.f; met[2] = s2c;
      } else {
        const float health = dm->healthy_reward * fabsf((site[2] < dm->endpoint_min_z_pos ? 1.f : 0.f) - 1.f);
        reward = push + s2c + health + site_z_reward;
        done = cube[2] < 0.6f ? 1.f : 0.f;
        met[0] = push; met[2] = s2c;
      }
    }
    reward = clipf(reward, -1e2f, 1e2f);
    if (kind == RSRX_ENV_SF) info[RSRX_INFO_LAST_ACTION] = last_action;
    get_obs(dm, sm, info);
    for (int i = 0; i < 3; i++) { info[RSRX_INFO_SITE + i] = site[i]; info[RSRX_INFO_OBJ + i] = sm[ar::XPOS + dm->cube_body * 3 + i]; }
    // EpisodeWrapper
    float steps = steps0 + (float)dm->action_repeat;
    float trunc = 0.f;
    if (dm->episode_length > 0 && steps >= (float)dm->episode_length) { trunc = 1.f - done; done = 1.f; }
    info[RSRX_INFO_STEPS] = steps;
    info[RSRX_INFO_TRUNCATION] = trunc;
    st.reward[e] = reward;
    st.done[e] = done;
  }
  done = __shfl_sync(0xffffffffu, done, 0);
  // non-finite guard
  bool bad = false;
#pragma unroll 1
  for (int i = lane; i < dm->nq; i += 32) bad |= !isfinite(sm[ar::QPOS + i]);
#pragma unroll 1
  for (int i = lane; i < dm->nv; i += 32) bad |= !isfinite(sm[ar::QVEL + i]);
  if (__any_sync(0xffffffffu, bad)) {
    status |= RSRX_STATUS_NONFINITE;
    // Deliberate deviation from MJX / brax, which would carry the NaN state (and NaN rewards and observations, i.e. NaN
    // gradients in a trainer) to the end of the episode: a wrapped env whose simulation blew up — an object knocked
    // off the table at tens of rad/s, ~1 env-step in 3e7 under uniform random actions — is terminated here with zero
    // reward and auto-reset like any other `done`, its info / metrics put back to their reset values; the sticky
    // status bit reports it.  The bare env (episode_length <= 0) keeps MJX's behaviour.
    if (dm->episode_length > 0) {
      done = 1.f;
      if (lane == 0) {
        const float* frow = st.first_data + (size_t)e * L.data_stride;
        st.reward[e] = 0.f;
        st.done[e] = 1.f;
        info[RSRX_INFO_TRUNCATION] = 0.f;
        info[RSRX_INFO_LAST_ACTION] = 0.f;
        for (int i = 0; i < 3; i++) {
          info[RSRX_INFO_SITE + i] = frow[L.site_xpos + dm->site_endpoint * 3 + i];
          info[RSRX_INFO_OBJ + i] = frow[L.xpos + dm->cube_body * 3 + i];
        }
        if (kind == RSRX_ENV_T) {
          info[RSRX_INFO_NEWPOS] = 0.24739072f; info[RSRX_INFO_NEWPOS + 1] = -0.00496255f;
          info[RSRX_INFO_XITA] = 0.2876f;
        } else {
          info[RSRX_INFO_NEWPOS] = 0.37342f; info[RSRX_INFO_NEWPOS + 1] = -0.07989f;
        }
        for (int i = 0; i < METRICS_STRIDE; i++) st.metrics[(size_t)e * METRICS_STRIDE + i] = 0.f;
      }
    }
  }
  if (lane == 0 && status) st.status[e] |= status;
  RSRX_SYNC();
  // ---- AutoReset post: pipeline_state and obs only (episode_length <= 0: bare env, no wrappers)
  if (done != 0.f && dm->episode_length > 0) {
    const float* frow = st.first_data + (size_t)e * L.data_stride;
#pragma unroll 1
    for (int i = lane; i < L.data_stride; i += 32) row[i] = frow[i];
#pragma unroll 1
    for (int i = lane; i < OBS_STRIDE; i += 32) st.obs[(size_t)e * OBS_STRIDE + i] = st.first_obs[(size_t)e * OBS_STRIDE + i];
  } else {
    store_env(dm, sm, lane, row, time);
#pragma unroll 1
    for (int i = lane; i < OBS_STRIDE; i += 32) st.obs[(size_t)e * OBS_STRIDE + i] = sm[ar::OBSBUF + i];
  }
  return true;
}

// raw nsteps x mjx.step on a data row (parity tests, friction sweep set-up); dump: debug internals of the last forward()
__device__ __forceinline__ bool physics_body(const DModel* __restrict__ dm, float* sm, int lane, int e, int spill_slot,
                                             float* __restrict__ data, int nsteps, const PerEnv& pe,
                                             int* __restrict__ status_out, float* __restrict__ dump, int status_init) {
  const rsrx_layout& L = dm->lay;
  float* row = data + (size_t)e * L.data_stride;
  load_env(dm, sm, lane, e, spill_slot, row, pe);
  float time = row[L.time];
  int status = status_init;
  SolverDims sd;
  sd.nsr = sd.ncon = sd.nrow = 0;
  int niter = 0;
  for (int f = 0; f < nsteps; ++f) {
    niter = forward<false>(dm, sm, lane, &sd, &status);
    if (dump) {
      float* dp = dump + (size_t)e * dbg::STRIDE;
      const int nv = dm->nv;
#pragma unroll 1
      for (int i = lane; i < nv * nv; i += 32) {
        const int r = i / nv, c = i % nv, hi = r > c ? r : c, lo = r > c ? c : r;
        dp[dbg::M + i] = sm[ar::MM + ((hi * (hi + 1)) >> 1) + lo];
      }
#pragma unroll 1
      for (int i = lane; i < nv; i += 32) {
        dp[dbg::BIAS + i] = sm[ar::V_BIAS + i];
        dp[dbg::QACC_SMOOTH + i] = sm[ar::V_QACCS + i];
        dp[dbg::QACC + i] = sm[ar::V_QACC + i];
        dp[dbg::QFRC_C + i] = sm[ar::V_QFRCC + i];
        dp[dbg::QFRC_ACT + i] = sm[ar::V_ACT + i];
      }
      if (lane == 0) { dp[dbg::NCON] = (float)sd.ncon; dp[dbg::NEFC] = (float)sd.nrow; dp[dbg::NITER] = (float)(niter & 0xff); dp[dbg::LS_TOTAL] = (float)(niter >> 8); }
      // the contact positions lived in storage the solver has reused: qpos is unchanged, so collision() rebuilds them
      {
        int scratch_status = 0;
        collision(dm, sm, lane, &scratch_status);
      }
      for (int c = lane; c < MAXC; c += 32) {
        const float* cr = sm + ar::CON + c * ar::CSTRIDE;
        const float* ctm = sm + ar::CTMP + c * ar::CTSTRIDE;
        const bool v = c < sd.ncon;
        dp[dbg::CDIST + c] = v ? cr[cf::DIST] : 0.f;
        for (int i = 0; i < 3; i++) dp[dbg::CPOS + c * 3 + i] = v ? ctm[ct::POS + i] : 0.f;
        const int bodies = __float_as_int(cr[cf::BODIES]);
        dp[dbg::CGEOM + c] = v ? (float)(((bodies >> 16) & 0xff) * 64 + ((bodies >> 24) & 0xff)) : -1.f;
      }
    }
    implicit_advance(dm, sm, lane);
    time += dm->timestep;
  }
  status = __reduce_or_sync(0xffffffffu, status);
  if ((status & RSRX_STATUS_CONTACT_OVERFLOW) && pe.redo && !dump) return false;
  store_env(dm, sm, lane, row, time);
  if (lane == 0 && status_out) status_out[e] |= status;
  return true;
}

#if !defined(RSRX_REDO_ONLY) && !defined(RSRX_STEP_ONLY)
// ---------------------------------------------------------------- reset kernel
__global__ void __launch_bounds__(32 * WPB) reset_kernel(const DModel* __restrict__ dm, int N, const float* __restrict__ qpos,
                                                  const float* __restrict__ qvel, const float* __restrict__ ctrl,
                                                  PerEnv pe, StatePtrs st) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, e = blockIdx.x * (int)(blockDim.x >> 5) + wib;
  float* sm = smem + wib * dm->arena_stride;
  if (e >= N) return;
  if (!env_reset_body(dm, sm, lane, e, e, qpos, qvel, ctrl, pe, st, 0) && lane == 0) {
    if (pe.redo) redo_push(pe, e); else st.status[e] = RSRX_STATUS_CONTACT_OVERFLOW;
  }
}

#endif
#ifndef RSRX_REDO_ONLY
// ----------------------------------------------------------------- step kernel
// MAXW: the most warps a CTA of this instantiation is launched with.  Up to 14 the register file allows 128 registers
// per thread; the 19-warp shape (3 rounds at 8192 envs) has to live with 96.
template <int MAXW>
__global__ void __launch_bounds__(32 * MAXW) step_kernel(const DModel* __restrict__ dm, int N, const float* __restrict__ action,
                                                 PerEnv pe, StatePtrs st) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, e = blockIdx.x * (int)(blockDim.x >> 5) + wib;
  float* sm = smem + wib * dm->arena_stride;
  if (e >= N) {
    for (int f = 0; f < dm->n_frames * kPhaseBarriers; ++f) phase_barrier<true>(__builtin_ctz(RSRX_SYNC_MASK));  // shadow the phase barriers
    return;
  }
  if (!env_step_body<true>(dm, sm, lane, e, e, action, pe, st, 0) && lane == 0) {
    if (pe.redo) redo_push(pe, e); else st.status[e] |= RSRX_STATUS_CONTACT_OVERFLOW;
  }
}

#endif
#if !defined(RSRX_REDO_ONLY) && !defined(RSRX_STEP_ONLY)
// ------------------------------------------------------------ physics-only kernel
__global__ void __launch_bounds__(32 * WPB) physics_kernel(const DModel* __restrict__ dm, int N, float* __restrict__ data,
                                                    int nsteps, PerEnv pe, int* __restrict__ status_out,
                                                    float* __restrict__ dump) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, e = blockIdx.x * (int)(blockDim.x >> 5) + wib;
  float* sm = smem + wib * dm->arena_stride;
  if (e >= N) return;
  if (!physics_body(dm, sm, lane, e, e, data, nsteps, pe, status_out, dump, 0) && lane == 0) redo_push(pe, e);
}
#endif  // RSRX_REDO_ONLY, RSRX_STEP_ONLY

#ifndef RSRX_STEP_ONLY
// ------------------------------------------------------------------ redo kernel
// Runs the env bodies for the envs on the redo list (persistent: warp w of the grid takes entries w, w + W, ...); the
// last CTA to finish clears the list for the next launch.  Meant for the large-capacity instantiation (rsrx_redo.cu),
// where contact_cap = 4 slots x every geom pair and a body cannot fail.
struct RedoArgs {
  int mode;  // 0 reset, 1 step, 2 physics
  const float *qpos, *qvel, *ctrl;  // reset
  const float* action;              // step
  float* data; int nsteps; int* status_out;  // physics
};
template <int W>
__global__ void __launch_bounds__(32 * W) redo_kernel(const DModel* __restrict__ dm, RedoArgs a, PerEnv pe, StatePtrs st) {
  extern __shared__ float smem[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, slot = blockIdx.x * W + wib;
  float* sm = smem + wib * dm->arena_stride;
  const int n = *reinterpret_cast<volatile int*>(pe.redo);
  PerEnv pq = pe;
  pq.redo = nullptr;
#pragma unroll 1
  for (int i = slot; i < n; i += gridDim.x * W) {
    const int e = pe.redo[2 + i];
    bool ok;
    if (a.mode == 0) ok = env_reset_body(dm, sm, lane, e, slot, a.qpos, a.qvel, a.ctrl, pq, st, RSRX_STATUS_CONTACT_REDO);
    else if (a.mode == 1) ok = env_step_body<false>(dm, sm, lane, e, slot, a.action, pq, st, RSRX_STATUS_CONTACT_REDO);
    else ok = physics_body(dm, sm, lane, e, slot, a.data, a.nsteps, pq, a.status_out, nullptr, RSRX_STATUS_CONTACT_REDO);
    if (!ok && lane == 0) st.status[e] |= RSRX_STATUS_CONTACT_OVERFLOW;  // unreachable while contact_cap >= 4 * npair
    RSRX_SYNC();
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    if (atomicAdd(pe.redo + 1, 1) == (int)gridDim.x - 1) {
      pe.redo[0] = 0;
      pe.redo[1] = 0;
      __threadfence();
    }
  }
}

#endif  // RSRX_STEP_ONLY
#if !defined(RSRX_REDO_ONLY) && !defined(RSRX_STEP_ONLY)
// Debug/parity entry: the cooperative narrow phase on a batch of explicit geom pairs (one half warp per pair, as in
// collision()).  in: [n][30] = p1(3) m1(9) s1(3) p2(3) m2(9) s2(3); plane != 0: geom 1 is a plane (s1 unused).
// out: [n][19] = dist(4) pos(4x3) nrm(3).
__global__ void __launch_bounds__(32) narrowphase_kernel(const float* __restrict__ in, int n, int plane, float* __restrict__ out) {
  __shared__ float scratch[2 * 24];
  __shared__ float pose[2][30];
  const int lane = threadIdx.x;
  Half hw;
  hw.shift = lane & 16; hw.l = lane & 15; hw.mask = 0xffffu << hw.shift;
  const int h = lane >> 4, pair = blockIdx.x * 2 + h;
  const bool valid = pair < n;
  for (int i = hw.l; i < 30; i += 16) pose[h][i] = in[(size_t)(valid ? pair : 0) * 30 + i];
  __syncwarp();
  if (valid) {
    const float* q = pose[h];
    const float s1[3] = {q[12], q[13], q[14]}, s2[3] = {q[27], q[28], q[29]};
    Manifold mf;
    if (plane) plane_box(q, q + 3, q + 15, q + 18, s2, hw, &mf);
    else box_box(q, q + 3, s1, q + 15, q + 18, s2, scratch + h * 24, hw, &mf);
    float* o = out + (size_t)pair * 19;
    if (hw.l < 4) {
      o[hw.l] = mf.dist;
      for (int i = 0; i < 3; i++) o[4 + hw.l * 3 + i] = mf.pos[i];
    }
    if (hw.l == 0) for (int i = 0; i < 3; i++) o[16 + i] = mf.nrm[i];
  }
}

#endif  // RSRX_REDO_ONLY

}  // namespace RSRX_NS
