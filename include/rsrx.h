/* rsrx.h — C-ABI of librsrx.so, the B200 (sm_100a) batched Airbot stepper.
 *
 * The reference has no native boundary on this path: callers go through the
 * brax `Env` API — `reset(rng) -> State`, `step(state, action) -> State`
 * (reference test/airbot.py:102,165; cube_env.py:95,145; T_shape_env.py:98,139)
 * wrapped by brax's VmapWrapper / DomainRandomizationVmapWrapper, EpisodeWrapper
 * and AutoResetWrapper (twin: ppo_train/go2_training/mujoco_playground/_src/
 * wrapper.py:117-165), and underneath `PipelineEnv.pipeline_init/pipeline_step`
 * -> `mjx.forward` / n_frames x `mjx.step` (twin call site:
 * mujoco_playground/_src/mjx_env.py:30-65).  The entry points below are what an
 * FFI for that path binds; INTEGRATION.md shows the ctypes and XLA-FFI stubs.
 *
 * Conventions
 *  - plain pointers and sizes; no torch / CUDA types in signatures.  `stream`
 *    is a `cudaStream_t` passed as `void*` (NULL = default stream).
 *  - every `float*` / `int32_t*` below is DEVICE memory owned by the caller,
 *    unless the name ends in `_host`.
 *  - functions return 0 on success, non-zero on error; rsrx_last_error() gives
 *    the thread-local message.  Nothing throws across the ABI.  Launches are
 *    asynchronous on `stream`; no host sync happens in rsrx_env_reset /
 *    rsrx_env_step, and the only allocation is the model's scratch buffer
 *    (5.4 KB per env), grown the first time a call sees a larger N: call reset
 *    once for a batch size before capturing steps in a CUDA graph.
 *  - batched natively: one call advances N environments; it *is* the
 *    AutoReset(Episode(Vmap|DomainRandomizationVmap(env))) stack.
 */
#ifndef RSRX_H_
#define RSRX_H_

#include <stddef.h>
#include <stdint.h>

#include "rsrx_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rsrx_model rsrx_model; /* opaque; holds the device copy of the model */

/* Offsets (in floats) of the fields inside one row of the `data` buffer — the
 * mjx.Data subset that callers of the Airbot envs read.
 * One row per env, `data_stride` floats apart. */
typedef struct rsrx_layout {
  int32_t data_stride; /* floats per env in data / first_data */
  int32_t qpos, qvel, ctrl, qacc_warmstart, time;
  int32_t xpos, xquat, site_xpos, geom_xpos;
  int32_t obs_size;    /* 23 (sf, cube) or 16 (T) */
  int32_t obs_stride;  /* floats per env in obs / first_obs */
  int32_t info_stride; /* floats per env in info (RSRX_INFO_*) */
  int32_t metrics_stride;
  int32_t nq, nv, nu, nbody, nsite, ngeom;
} rsrx_layout;

/* info row (floats).  sf/cube: target_pos, new_cube_pos, site_pos, cube_pos,
 * last_action.  T: target_base_pos, target_vertical_pos, target_w, new_T_pos,
 * site_pos, T_pos, xita.  (reference test/airbot.py:153-159, T_shape_env.py:125-132) */
#define RSRX_INFO_TARGET 0   /* [3] target_pos | target_base_pos */
#define RSRX_INFO_TARGET2 3  /* [3] T: target_vertical_pos */
#define RSRX_INFO_NEWPOS 6   /* [2] new_cube_pos | new_T_pos */
#define RSRX_INFO_SITE 8     /* [3] site_pos */
#define RSRX_INFO_OBJ 11     /* [3] cube_pos | T_pos */
#define RSRX_INFO_LAST_ACTION 14
#define RSRX_INFO_XITA 15
#define RSRX_INFO_TARGET_W 16
#define RSRX_INFO_STEPS 17      /* brax EpisodeWrapper info['steps'] */
#define RSRX_INFO_TRUNCATION 18 /* brax EpisodeWrapper info['truncation'] */
#define RSRX_INFO_STRIDE 20

/* per-env status bits written by the kernels (no host sync needed to keep going) */
#define RSRX_STATUS_NONFINITE 1      /* a non-finite value reached qpos/qvel.  rsrx_env_step on a wrapped env
                                      * (episode_length > 0) then ends the env's episode in that step: done = 1,
                                      * reward 0, state / obs auto-reset, info / metrics back to their reset values
                                      * (MJX would carry the NaN to the end of the episode); sticky, informational */
#define RSRX_STATUS_CONTACT_OVERFLOW 2 /* contacts were dropped.  Cannot happen in reset / step / physics_step: an env
                                        * whose substep has more active contacts than the fast kernel's arena holds
                                        * (rsrx_max_contacts()) is re-run by the large-capacity kernel, which keeps every
                                        * slot of every geom pair (4 x npair) like MJX.  Only the debug dump can set it. */
#define RSRX_STATUS_SOLVER_CAP 4     /* Newton hit opt.iterations */
#define RSRX_STATUS_CONTACT_REDO 8   /* informational: at least one step of this env went through the large-capacity kernel */

/* State of N wrapped envs; every pointer is a caller-owned device buffer. */
typedef struct rsrx_state {
  float* data;        /* [N][data_stride]  pipeline_state                 */
  float* first_data;  /* [N][data_stride]  info['first_pipeline_state']   */
  float* obs;         /* [N][obs_stride]                                   */
  float* first_obs;   /* [N][obs_stride]   info['first_obs']               */
  float* reward;      /* [N] */
  float* done;        /* [N] */
  float* info;        /* [N][RSRX_INFO_STRIDE] */
  float* metrics;     /* [N][metrics_stride] */
  int32_t* status;    /* [N] RSRX_STATUS_* (OR-accumulated) */
} rsrx_state;

/* Per-env model arrays = the four leaves the reference's domain_randomize gives a
 * leading env axis (ppo_train/airbot_training/domain_randomize.py:71-90) — also
 * used for the friction sweep (RSR/rsr_pipeline.py:125-136 sets
 * geom_friction[-1,:]).  NULL = the nominal model value. */
typedef struct rsrx_per_env {
  const float* geom_friction;    /* [N][ngeom][3] */
  const float* body_mass;        /* [N][nbody]    */
  const float* dof_damping;      /* [N][nv]       */
  const float* dof_frictionloss; /* [N][nv]       */
} rsrx_per_env;

/* ---- model ----------------------------------------------------------------
 * replaces mujoco.MjModel.from_xml_path + brax.io.mjcf.load_model + mjx.put_model
 * (test/airbot.py:43-47): the host compiles MJCF to an rsrx_model_blob
 * (rsr_mjx_b200/mjcf.py) and hands it over here together with the env constants. */
int rsrx_model_create(const void* blob_host, size_t blob_bytes, const rsrx_env_cfg* cfg_host, rsrx_model** out);
void rsrx_model_destroy(rsrx_model* m);
int rsrx_model_layout(const rsrx_model* m, rsrx_layout* out);
size_t rsrx_model_blob_size(void);
size_t rsrx_env_cfg_size(void);

/* ---- env.reset -------------------------------------------------------------
 * replaces jit(vmap(env.reset))(keys) (RSR/train.py:231) minus the sampling,
 * which stays host-side Python (rsr_mjx_b200/airbot_spec.py::sample_reset restates
 * jax.random): given sampled qpos[N][nq], qvel[N][nv], ctrl[N][nu] it runs
 * pipeline_init (= mjx.make_data + mjx.forward, mjx_env.py:30-54), sets ctrl,
 * builds info / metrics / obs (test/airbot.py:135-163) and the wrappers' reset
 * state (first_pipeline_state, first_obs, steps, truncation). */
int rsrx_env_reset(const rsrx_model* m, int N, const float* qpos, const float* qvel, const float* ctrl,
                   const rsrx_per_env* per_env, rsrx_state st, void* stream);

/* ---- env.step --------------------------------------------------------------
 * replaces AutoResetWrapper(EpisodeWrapper(VmapWrapper(env))).step(state, action)
 * (RSR/train.py:313 actor_step -> env.step): action[N][nu] in [-1,1]; one launch
 * does action shaping, n_frames x mjx.step, reward/obs/done, episode bookkeeping
 * and auto-reset for all N envs, in place on `st`. */
int rsrx_env_step(const rsrx_model* m, int N, rsrx_state st, const float* action, const rsrx_per_env* per_env,
                  void* stream);

/* Same step for a caller whose policy lives on the host: host_action[N][nu] (pinned memory recommended) is copied to
 * the device buffer `action_staging` [N][nu], the step runs, and obs [N][obs_stride] / reward [N] / done [N] are copied
 * back into the host buffers that are not NULL — all asynchronously on `stream`; synchronise the stream before reading
 * them.  One call per step instead of copy + launch + gather + copy. */
int rsrx_env_step_host(const rsrx_model* m, int N, rsrx_state st, const float* host_action, float* action_staging,
                       float* host_obs, float* host_reward, float* host_done, const rsrx_per_env* per_env, void* stream);

/* ---- physics only (pipeline_step / mjx.step, mjx_env.py:55-65) ---------------
 * advances data rows by nsteps x mjx.step with the ctrl stored in the rows. */
int rsrx_physics_step(const rsrx_model* m, int N, float* data, int nsteps, const rsrx_per_env* per_env,
                      int32_t* status, void* stream);

/* Debug dump of the last forward() inside rsrx_physics_step(nsteps=1) for parity
 * tests: per env M[nv*nv], qfrc_bias[nv], qacc_smooth[nv], qacc[nv],
 * qfrc_constraint[nv], ncon, nefc, niter, contact dist[cap], contact pos[cap*3].
 * `dump` is [N][rsrx_debug_stride()] floats. */
int rsrx_debug_stride(void);
int rsrx_max_contacts(void); /* active contacts per env the fast kernel holds (beyond it: RSRX_STATUS_CONTACT_REDO) */
int rsrx_physics_step_debug(const rsrx_model* m, int N, float* data, const rsrx_per_env* per_env, float* dump,
                            void* stream);

/* ---- RSR distribution loss (RSR/rsr_loss.py:122-175, dataset_processor.py:17-43)
 * density[M] = softmax_m(logsumexp_n(-|grid_m - x_n|^2 / (2 h^2)) - log Ntot) over
 * x = [reference_data (Nref rows); online batch (Nb rows)], both [.,D] row-major;
 * distance = sum|cumsum(density) - cumsum(reference_density)|;
 * loss = loss_scale * divergence * distance.
 * out_host-free: out[0]=loss, out[1]=distance (device).  grad_batch may be NULL;
 * otherwise it receives d loss / d batch [Nb][D] (callers slice the action
 * columns, the only ones a policy gradient flows through).  Two launches on
 * `stream` (KDE + tail, gradient) sharing a library-owned per-device scratch
 * ((Nref + Nb) * M floats, grown the first time a larger size is seen: call once
 * before capturing a CUDA graph; calls on one device must be stream-ordered). */
int rsrx_rsr_loss(const float* grid, int M, int D, const float* reference_data, int Nref, const float* batch, int Nb,
                  const float* reference_density, float bandwidth, float divergence, float loss_scale,
                  float* density_out, float* out, float* grad_batch, void* stream);
/* The RSR term as the PPO loss uses it (RSR/losses.py:186-195: compute_rsr_loss on (observation, mode of the policy,
 * next_observation)) without host-framework glue.  rsrx_rsr_policy_term: packs transition[r] = [obs[r] (O) |
 * tanh(logits[r][0:A]) | next_obs[r] (O)] into `transition` [rows][2O+A] and runs rsrx_rsr_loss on it: out[0] = loss,
 * out[1] = distance, grad_transition [rows][2O+A] = d loss / d transition.  rsrx_rsr_logit_grad: grad_logits_out[r][k] =
 * grad_logits_in[r][k] (NULL = 0) + d loss / d logits[r][k] (chain rule through tanh for k < A, zero for the scale half);
 * in and out may alias.  logits are [rows][2A] (location | scale parameters). */
int rsrx_rsr_policy_term(const float* grid, int M, const float* reference_data, int Nref, const float* reference_density,
                         float bandwidth, float divergence, float loss_scale, const float* obs, const float* logits,
                         const float* next_obs, int rows, int O, int A, float* transition, float* grad_transition, float* out,
                         void* stream);
int rsrx_rsr_logit_grad(const float* transition, const float* grad_transition, const float* grad_logits_in, int rows, int O, int A,
                        float* grad_logits_out, void* stream);
/* KDE density only (evaluate_kde) for build_rsr_data (rsr_loss.py:43-91) */
int rsrx_kde(const float* grid, int M, int D, const float* data, int Ndata, float bandwidth, float* density_out,
             void* stream);

/* Parity/debug: the narrow phase of mjx collision_convex (box_box; plane != 0: plane_convex on a box) on n explicit geom
 * pairs.  pairs [n][30] = pos1(3) mat1(9, row-major) size1(3) pos2(3) mat2(9) size2(3); out [n][19] = dist(4)
 * pos(4x3) normal(3), a slot with dist >= 0 holds no contact.  Device pointers. */
int rsrx_debug_narrowphase(const float* pairs, int n, int plane, float* out, void* stream);

/* Behaviour-policy head of the actor step (brax NormalTanhDistribution, min_std 0.001): logits [N][2A] (loc |
 * pre-softplus scale) and noise [N][A] ~ N(0,1) give raw_action = loc + scale * noise, action = tanh(raw_action) and
 * log_prob [N] of raw_action under the tanh-normal (NULL to skip).  Device float32 arrays, one launch. */
int rsrx_tanh_normal_act(const float* logits, const float* noise, int N, int A, float* raw_action, float* action,
                         float* log_prob, void* stream);

/* Minibatch gather for the trainers: dst[k][r][:] = src[k][idx[r]][:] for nfields (<= 8) row-major float tensors with
 * row_floats[k] floats per row, one launch.  src / dst / row_floats are HOST arrays (of device pointers / sizes), idx is a
 * device array of nrows int64 row indices. */
int rsrx_gather_rows(const float* const* src, float* const* dst, const int32_t* row_floats, int nfields,
                     const int64_t* idx, int nrows, void* stream);

/* MLP backward helper for the trainers' networks: grad_z = grad_y * act'(z) (activation 0 none, 1 silu, 2 relu;
 * grad_z may be NULL when activation is 0) and grad_bias[c] = sum over rows of grad_z[r][c], one deterministic launch.
 * Row-major [rows][cols] device arrays; workspace: rsrx_act_bias_backward_workspace(rows, cols) floats whose FIRST
 * word is zero before the first call (the kernel leaves it zero). */
size_t rsrx_act_bias_backward_workspace(int rows, int cols);
int rsrx_act_bias_backward(const float* grad_y, const float* z, int rows, int cols, int activation, float* grad_z,
                           float* grad_bias, float* workspace, void* stream);

/* Fused PPO loss head, forward + backward in one launch (RSR/losses.py:39-95 compute_gae, :98-205 compute_ppo_loss
 * between the network outputs and the scalar task loss; brax NormalTanhDistribution with min_std 0.001).
 * All arrays are device float32, batch-major: logits [B][T][2A] (loc | pre-softplus scale), baseline / behaviour_log_prob /
 * reward / discount / truncation [B][T], bootstrap_value [B] (value of the last next_observation), raw_action / noise
 * [B][T][A] (pre-tanh behaviour action; N(0,1) sample for the entropy estimate).  workspace: 2*B*T floats.
 * out[4] = task_loss (= policy + value + entropy), policy_loss, v_loss, entropy_loss;
 * grad_logits [B][T][2A] and grad_baseline [B][T] receive d task_loss / d (logits, baseline); vs and the advantages are
 * stop-gradient as in the reference, so bootstrap_value has no gradient. */
int rsrx_ppo_head(const float* logits, const float* baseline, const float* bootstrap_value, const float* raw_action,
                  const float* behaviour_log_prob, const float* reward, const float* discount, const float* truncation,
                  const float* noise, int B, int T, int A, float reward_scaling, float discounting, float gae_lambda,
                  float clipping_epsilon, float entropy_cost, int normalize_advantage, float* workspace, float* out,
                  float* grad_logits, float* grad_baseline, void* stream);

/* Minibatch input preparation of the fused PPO update (RSR/losses.py:120-131: normalize_fn(data.observation), the
 * bootstrap observation data.next_observation[-1]) in one launch: obs / next_obs [mb][T][O], running-statistics mean / std
 * [O] -> obs_n [mb*T][O] (policy input), x_pad [mb*T + mb][ldp] (value input: the normalised observations then the mb
 * normalised bootstrap observations, zero-padded to ldp >= O columns) and its transpose xT [ldp][ldt >= mb*T + mb]. */
int rsrx_ppo_prep(const float* obs, const float* next_obs, const float* mean, const float* std, int mb, int T, int O,
                  float* obs_n, float* x_pad, int ldp, float* xT, int ldt, void* stream);

/* ---- tensor-core linear layers for the trainers' value networks (csrc/rsrx_gemm.cuh: TMA loads, tcgen05.mma kind::tf32,
 * fp32 accumulation in TMEM, fused epilogues).  Replaces torch.addmm + SiLU + SiLU' + bias-gradient launches of the value
 * MLP of RSR/train.py (brax make_ppo_networks, value_hidden_layer_sizes (256,)*5) / the critics of RSR/sac_train.py.
 * Row-major device float32 arrays; every pointer 16-byte aligned, every leading dimension a multiple of 4.
 * activation: 0 none, 1 silu (brax swish), 2 relu.  Operands that are contiguous along the contraction are loaded by TMA
 * (cp.async.bulk.tensor, 128-byte swizzle); the others are transposed by the threads while being staged (slower), which
 * is why the epilogues can also write TRANSPOSED copies (yT / dzprevT: [N][ldt >= M]) for the weight-gradient GEMM.
 *   forward : z[M][N] = x[M][K] w[N][K]^T + bias (z may be NULL), y = act(z), optional yT
 *   dgrad   : dzprev[M][Nin] = (dz[M][Nout] w[Nout][Nin]) * act'(zprev); pass wT [Nin][ldwt >= Nout] (a transposed copy
 *             of w, e.g. kept by rsrx_adam_step) to use TMA for it, else w; if colsum_partials != NULL the column sums of
 *             dzprev over each block of 128 rows -> colsum_partials[ceil(M/128)][ld] (the previous layer's bias gradient,
 *             finished by rsrx_reduce_partials); optional dzprevT
 *   wgrad   : partials[s][Nout][ldp] = dz[rows_s][Nout]^T x[rows_s][Nin] for row slices of rows_per_split rows
 *             (ceil(rows / rows_per_split) slices); summed in slice order by rsrx_reduce_partials: deterministic.
 *             transposed_inputs != 0: dz / x are the transposed copies dzT [Nout][lddz >= rows], xT [Nin][ldx >= rows]
 *   reduce  : out_k[i] = sum_{s < S_k} in_k[s * stride_k + i], i < n_k, for up to 24 segments in one launch; in / out /
 *             n / S / stride are HOST arrays
 *   value_head_backward: the scalar output layer v = h . w + b fused with the last hidden layer's activation derivative:
 *             dz[m][j] = g[m] w[j] act'(z[m][j]) (+ optional dzT) and per-128-row partials of colsum(dz), dw[j] = sum g[m]
 *             h[m][j], db = sum g[m] (colsum_partials [blocks][ld], dw_partials [blocks][n], db_partials [blocks]) */
int rsrx_linear_forward(const float* x, int ldx, const float* w, int ldw, const float* bias, int M, int N, int K,
                        int activation, float* z, float* y, int ldy, float* yT, int ldt, void* stream);
int rsrx_linear_dgrad(const float* dz, int lddz, const float* w, int ldw, const float* wT, int ldwt, const float* zprev, int M,
                      int Nin, int Nout, int activation, float* dzprev, int ld, float* colsum_partials, float* dzprevT, int ldt,
                      void* stream);
int rsrx_linear_wgrad(const float* dz, int lddz, const float* x, int ldx, int transposed_inputs, int rows, int Nout, int Nin,
                      int rows_per_split, float* partials, int ldp, void* stream);
int rsrx_reduce_partials(const float* const* in, float* const* out, const int32_t* n, const int32_t* S,
                         const int64_t* stride, int nseg, void* stream);
int rsrx_value_head_backward(const float* g, const float* w, const float* z, const float* h, int M, int n, int ld,
                             int activation, float* dz, float* colsum_partials, float* dw_partials, float* db_partials,
                             float* dzT, int ldt, void* stream);

/* ---- the trainers' policy network (brax make_ppo_networks policy_hidden_layer_sizes (32,)*4; csrc/rsrx_mlp.cuh):
 * an MLP with every width <= 32 and <= 8 layers, hidden activation 1 silu / 2 relu, linear output, one warp per row.
 * weights[l] is [widths[l+1]][widths[l]] row-major (torch nn.Linear), biases[l] [widths[l+1]]; weights / biases / widths
 * are HOST arrays (device pointers / ints).
 *   forward : out[rows][ldo] = MLP(x[rows][ldx]); zs [nlayers-1][rows][32] receives the hidden pre-activations;
 *             norm_mean / norm_std (both or neither, [widths[0]]): the input is normalised on the way in,
 *             (x - mean) / std — the actor step's running-statistics normaliser (RSR/train.py:313) without a launch
 *   backward: from grad_out [rows][ldg] and zs, one partial gradient vector per CTA in parameter order (W_0, b_0, W_1,
 *             b_1, ...): partials [rsrx_small_mlp_backward_ctas(rows)][total]; finish with rsrx_reduce_partials */
int rsrx_small_mlp_forward(const float* const* weights, const float* const* biases, const int32_t* widths, int nlayers,
                           int activation, const float* x, int ldx, int rows, float* zs, float* out, int ldo,
                           const float* norm_mean, const float* norm_std, void* stream);
int rsrx_small_mlp_backward(const float* const* weights, const float* const* biases, const int32_t* widths, int nlayers,
                            int activation, const float* x, int ldx, int rows, const float* zs, const float* grad_out, int ldg,
                            float* partials, void* stream);
int rsrx_small_mlp_backward_ctas(int rows);

/* Adam (torch.optim.Adam / optax.adam semantics: bias-corrected moments, no weight decay) on up to 32 tensors in ONE
 * launch — the optimiser step of RSR/train.py:244-262 (optax.adam(learning_rate)) for the trainers' small networks.
 * params / grads / exp_avg / exp_avg_sq / sizes are HOST arrays of device pointers / element counts; the gradient is read
 * as grads[k][i] * grad_scale (1 / world_size after a sum all-reduce).  params_t / cols (both may be NULL): where
 * params_t[k] != NULL the updated tensor, seen as [sizes[k] / cols[k]][cols[k]], is also written transposed to
 * params_t[k] ([cols][rows]: the weight copy rsrx_linear_dgrad loads by TMA).  step_ticket: one device uint64, zero
 * before the first step, advanced by the kernel itself (so a captured CUDA graph keeps counting on every replay): its low
 * 32 bits are the number of steps taken, the high 32 bits a scratch counter that is zero between launches. */
int rsrx_adam_step(float* const* params, const float* const* grads, float* const* exp_avg, float* const* exp_avg_sq,
                   const int32_t* sizes, float* const* params_t, const int32_t* cols, int ntensors, float lr, float beta1,
                   float beta2, float eps, float grad_scale, uint64_t* step_ticket, void* stream);

const char* rsrx_last_error(void);
const char* rsrx_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RSRX_H_ */
