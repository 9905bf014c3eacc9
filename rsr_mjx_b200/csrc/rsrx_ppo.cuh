// rsrx_ppo.cuh — fused PPO loss head (forward + backward), one launch.
//   reference: RSR/losses.py:39-95 (compute_gae), :98-205 (compute_ppo_loss), brax NormalTanhDistribution
// Everything between the network outputs and the scalar loss in one CTA:
//   GAE (one thread per sequence, backward scan over T), advantage normalisation (two-pass mean / population std),
//   tanh-normal log-prob of the behaviour action, clipped surrogate, value loss against the (stop-gradient) vs,
//   entropy estimate from one noise sample, and the gradients w.r.t. policy logits and baseline values.
// All tensors are batch-major [B][T]...; B*T ~ 10^3..10^4, so this is latency-bound: it replaces ~150 elementwise
// launches of the eager/graph path (1.3 ms per minibatch step -> see profiles/).
#pragma once
#include <cuda_runtime.h>
#include <cooperative_groups.h>
#include <math.h>

#include <cstdlib>

#include "rsrx_pdl.cuh"

namespace rsrx {
namespace ppo {

constexpr int THREADS = 1024;
constexpr int MAXA = 16;
constexpr float MIN_STD = 0.001f;
constexpr float LOG_2PI = 1.8378770664093453f;
constexpr float LOG_2 = 0.6931471805599453f;

__device__ __forceinline__ float softplus(float x) { return x > 20.f ? x : log1pf(expf(x)); }  // torch threshold
__device__ __forceinline__ float sigmoid(float x) { return 1.f / (1.f + expf(-x)); }
// log |d tanh(x) / dx| = 2 (log 2 - x - softplus(-2x))
__device__ __forceinline__ float log_det_jac(float x) { return 2.f * (LOG_2 - x - softplus(-2.f * x)); }

// sum over the CTA; every thread gets the result.  red: THREADS/32 floats of shared memory
__device__ __forceinline__ float block_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
  for (int w = 0; w < THREADS / 32; w++) s += red[w];  // same order in every thread: identical bits
  return s;
}

struct Hyper {
  float reward_scaling, discounting, gae_lambda, clip_eps, entropy_cost;
  int normalize_advantage;
};

__global__ void __launch_bounds__(THREADS) head_kernel(
    const float* __restrict__ logits,      // [B][T][2A]
    const float* __restrict__ baseline,    // [B][T]
    const float* __restrict__ bootstrap,   // [B]
    const float* __restrict__ raw_action,  // [B][T][A]
    const float* __restrict__ behaviour_lp,  // [B][T]
    const float* __restrict__ reward, const float* __restrict__ discount, const float* __restrict__ truncation,  // [B][T]
    const float* __restrict__ noise,       // [B][T][A]
    int B, int T, int A, Hyper h,
    float* __restrict__ ws,                // workspace [2][B][T]: advantages, vs
    float* __restrict__ out,               // task_loss, policy_loss, v_loss, entropy_loss
    float* __restrict__ grad_logits,       // [B][T][2A]  d task_loss / d logits
    float* __restrict__ grad_baseline) {   // [B][T]
  __shared__ float red[THREADS / 32];
  pdl::launch_dependents();
  pdl::wait();
  const int n = B * T, tid = threadIdx.x;
  float* adv = ws;
  float* vs = ws + n;
  // ---- GAE: thread per sequence
  float s1 = 0.f;
  for (int b = tid; b < B; b += THREADS) {
    float acc = 0.f, vs_next = bootstrap[b], v_next = bootstrap[b];
    for (int t = T - 1; t >= 0; --t) {
      const int i = b * T + t;
      const float trunc = truncation[i], mask = 1.f - trunc;
      const float term = (1.f - discount[i]) * (1.f - trunc);
      const float r = reward[i] * h.reward_scaling, v = baseline[i];
      const float delta = (r + h.discounting * (1.f - term) * v_next - v) * mask;
      acc = delta + h.discounting * (1.f - term) * mask * h.gae_lambda * acc;
      const float vs_t = acc + v;
      const float a = (r + h.discounting * (1.f - term) * vs_next - v) * mask;
      adv[i] = a;
      vs[i] = vs_t;
      s1 += a;
      vs_next = vs_t;
      v_next = v;
    }
  }
  float mean = 0.f, inv_std = 1.f;
  if (h.normalize_advantage) {
    mean = block_sum(s1, red) / (float)n;  // (block_sum syncs: adv/vs are visible afterwards)
    float s2 = 0.f;
    for (int i = tid; i < n; i += THREADS) { const float d = adv[i] - mean; s2 += d * d; }
    inv_std = 1.f / (sqrtf(block_sum(s2, red) / (float)n) + 1e-8f);
  } else {
    __syncthreads();
  }
  // ---- per-transition losses and gradients
  const float inv_n = 1.f / (float)n;
  float l_pol = 0.f, l_v = 0.f, l_ent = 0.f;
  for (int i = tid; i < n; i += THREADS) {
    const float* lg = logits + (size_t)i * 2 * A;
    const float* ra = raw_action + (size_t)i * A;
    const float* nz = noise + (size_t)i * A;
    float lp = 0.f, ent = 0.f;
    float scale[MAXA], z[MAXA], sg[MAXA], dldj[MAXA];
#pragma unroll 1
    for (int k = 0; k < A; k++) {
      const float loc = lg[k], s = lg[A + k];
      const float sc = softplus(s) + MIN_STD;
      const float zz = (ra[k] - loc) / sc;
      lp += -0.5f * zz * zz - 0.5f * LOG_2PI - logf(sc) - log_det_jac(ra[k]);
      const float x = loc + sc * nz[k];
      ent += 0.5f + 0.5f * LOG_2PI + logf(sc) + log_det_jac(x);
      scale[k] = sc; z[k] = zz; sg[k] = sigmoid(s); dldj[k] = -2.f * tanhf(x);
    }
    const float a = (adv[i] - mean) * inv_std;
    const float rho = expf(lp - behaviour_lp[i]);
    const float lo = 1.f - h.clip_eps, hi = 1.f + h.clip_eps;
    const bool in_range = rho >= lo && rho <= hi;
    const float l1 = rho * a, l2 = fminf(fmaxf(rho, lo), hi) * a;
    l_pol += fminf(l1, l2);
    // d min(l1, l2) / d rho: l1's slope while l1 is the minimum (or both coincide inside the clip range)
    const float dmin_drho = (in_range || l1 < l2) ? a : 0.f;
    const float g_lp = -inv_n * dmin_drho * rho;              // d policy_loss / d lp
    const float g_ent = -h.entropy_cost * inv_n;              // d entropy_loss / d ent
    const float verr = vs[i] - baseline[i];
    l_v += verr * verr;
    l_ent += ent;
    grad_baseline[i] = -0.5f * verr * inv_n;
    float* gl = grad_logits + (size_t)i * 2 * A;
#pragma unroll 1
    for (int k = 0; k < A; k++) {
      const float d_loc = g_lp * (z[k] / scale[k]) + g_ent * dldj[k];
      const float d_scale = g_lp * ((z[k] * z[k] - 1.f) / scale[k]) + g_ent * (1.f / scale[k] + dldj[k] * nz[k]);
      gl[k] = d_loc;
      gl[A + k] = d_scale * sg[k];
    }
  }
  const float pol = -block_sum(l_pol, red) * inv_n;
  const float vl = block_sum(l_v, red) * inv_n * 0.25f;
  const float en = -h.entropy_cost * block_sum(l_ent, red) * inv_n;
  if (tid == 0) { out[0] = pol + vl + en; out[1] = pol; out[2] = vl; out[3] = en; }
}

// ---- the same head on a thread-block cluster of 8 CTAs -----------------------------------------------------------------
// The single-CTA kernel above is a 35 us serial section in the middle of every minibatch step (147 SMs idle).  Here CTA r
// of the cluster owns sequences [r B/8, (r+1) B/8): it stages their reward / value / discount / mask in shared memory
// with coalesced loads, runs the GAE recurrences from there, and handles their transitions; the five sums over all
// transitions (advantage mean and variance, the three losses) are CTA partials exchanged through distributed shared
// memory and added in rank order by every CTA (identical bits everywhere, deterministic).
namespace cg = cooperative_groups;
constexpr int CL = 8;     // CTAs per cluster (the portable maximum)
constexpr int CT = 512;   // threads per CTA

__device__ __forceinline__ float cta_sum(float v, float* red) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float s = 0.f;
#pragma unroll
  for (int w = 0; w < CT / 32; w++) s += red[w];
  return s;
}
// part: [8] slots of this CTA's shared memory, one per reduction (never reused within a launch)
__device__ __forceinline__ float cluster_sum(cg::cluster_group& cl, float v, float* red, float* part, int slot) {
  const float s = cta_sum(v, red);
  if (threadIdx.x == 0) part[slot] = s;
  cl.sync();
  float t = 0.f;
#pragma unroll
  for (int r = 0; r < CL; r++) t += cl.map_shared_rank(part, r)[slot];
  return t;
}

__global__ void __launch_bounds__(CT) head_cluster_kernel(
    const float* __restrict__ logits, const float* __restrict__ baseline, const float* __restrict__ bootstrap,
    const float* __restrict__ raw_action, const float* __restrict__ behaviour_lp, const float* __restrict__ reward,
    const float* __restrict__ discount, const float* __restrict__ truncation, const float* __restrict__ noise, int B, int T,
    int A, Hyper h, int cap, float* __restrict__ out, float* __restrict__ grad_logits, float* __restrict__ grad_baseline) {
  extern __shared__ float dyn[];  // [4][cap]: reward -> advantage, baseline, discount factor -> vs, mask
  __shared__ float red[CT / 32];
  __shared__ float part[8];
  cg::cluster_group cl = cg::this_cluster();
  pdl::launch_dependents();
  pdl::wait();
  const int tid = threadIdx.x, rank = (int)cl.block_rank();
  const int bper = (B + CL - 1) / CL, b0 = min(B, rank * bper), b1 = min(B, b0 + bper);
  const int i0 = b0 * T, nloc = (b1 - b0) * T, n = B * T;
  float* s_r = dyn;
  float* s_v = dyn + cap;
  float* s_c = dyn + 2 * cap;
  float* s_m = dyn + 3 * cap;
  for (int j = tid; j < nloc; j += CT) {
    const int i = i0 + j;
    const float trunc = truncation[i];
    const float term = (1.f - discount[i]) * (1.f - trunc);
    s_r[j] = reward[i] * h.reward_scaling;
    s_v[j] = baseline[i];
    s_c[j] = h.discounting * (1.f - term);
    s_m[j] = 1.f - trunc;
  }
  __syncthreads();
  // ---- GAE: thread per sequence, from shared memory (same arithmetic, in the same order, as head_kernel)
  float s1 = 0.f;
  for (int b = b0 + tid; b < b1; b += CT) {
    float acc = 0.f, vs_next = bootstrap[b], v_next = vs_next;
    for (int t = T - 1; t >= 0; --t) {
      const int j = (b - b0) * T + t;
      const float r = s_r[j], v = s_v[j], c = s_c[j], mask = s_m[j];
      const float delta = (r + c * v_next - v) * mask;
      acc = delta + c * mask * h.gae_lambda * acc;
      const float vs_t = acc + v;
      const float a = (r + c * vs_next - v) * mask;
      s_r[j] = a;     // advantage
      s_c[j] = vs_t;  // value target
      s1 += a;
      vs_next = vs_t;
      v_next = v;
    }
  }
  float mean = 0.f, inv_std = 1.f;
  if (h.normalize_advantage) {
    mean = cluster_sum(cl, s1, red, part, 0) / (float)n;
    float s2 = 0.f;
    for (int j = tid; j < nloc; j += CT) { const float d = s_r[j] - mean; s2 += d * d; }
    inv_std = 1.f / (sqrtf(cluster_sum(cl, s2, red, part, 1) / (float)n) + 1e-8f);
  } else {
    __syncthreads();
  }
  // ---- per-transition losses and gradients
  const float inv_n = 1.f / (float)n;
  float l_pol = 0.f, l_v = 0.f, l_ent = 0.f;
  for (int j = tid; j < nloc; j += CT) {
    const int i = i0 + j;
    const float* lg = logits + (size_t)i * 2 * A;
    const float* ra = raw_action + (size_t)i * A;
    const float* nz = noise + (size_t)i * A;
    float lp = 0.f, ent = 0.f;
    float scale[MAXA], z[MAXA], sg[MAXA], dldj[MAXA];
#pragma unroll 1
    for (int k = 0; k < A; k++) {
      const float loc = lg[k], s = lg[A + k];
      const float sc = softplus(s) + MIN_STD;
      const float zz = (ra[k] - loc) / sc;
      lp += -0.5f * zz * zz - 0.5f * LOG_2PI - logf(sc) - log_det_jac(ra[k]);
      const float x = loc + sc * nz[k];
      ent += 0.5f + 0.5f * LOG_2PI + logf(sc) + log_det_jac(x);
      scale[k] = sc; z[k] = zz; sg[k] = sigmoid(s); dldj[k] = -2.f * tanhf(x);
    }
    const float a = (s_r[j] - mean) * inv_std;
    const float rho = expf(lp - behaviour_lp[i]);
    const float lo = 1.f - h.clip_eps, hi = 1.f + h.clip_eps;
    const bool in_range = rho >= lo && rho <= hi;
    const float l1 = rho * a, l2 = fminf(fmaxf(rho, lo), hi) * a;
    l_pol += fminf(l1, l2);
    const float dmin_drho = (in_range || l1 < l2) ? a : 0.f;
    const float g_lp = -inv_n * dmin_drho * rho;
    const float g_ent = -h.entropy_cost * inv_n;
    const float verr = s_c[j] - s_v[j];
    l_v += verr * verr;
    l_ent += ent;
    grad_baseline[i] = -0.5f * verr * inv_n;
    float* gl = grad_logits + (size_t)i * 2 * A;
#pragma unroll 1
    for (int k = 0; k < A; k++) {
      const float d_loc = g_lp * (z[k] / scale[k]) + g_ent * dldj[k];
      const float d_scale = g_lp * ((z[k] * z[k] - 1.f) / scale[k]) + g_ent * (1.f / scale[k] + dldj[k] * nz[k]);
      gl[k] = d_loc;
      gl[A + k] = d_scale * sg[k];
    }
  }
  // the three loss sums: one exchange (slots 2..4), one cluster barrier
  const float p0 = cta_sum(l_pol, red), p1 = cta_sum(l_v, red), p2 = cta_sum(l_ent, red);
  if (tid == 0) { part[2] = p0; part[3] = p1; part[4] = p2; }
  cl.sync();
  if (rank == 0 && tid == 0) {
    float t0 = 0.f, t1 = 0.f, t2 = 0.f;
    for (int r = 0; r < CL; r++) {
      const float* q = cl.map_shared_rank(part, r);
      t0 += q[2]; t1 += q[3]; t2 += q[4];
    }
    const float pol = -t0 * inv_n, vl = t1 * inv_n * 0.25f, en = -h.entropy_cost * t2 * inv_n;
    out[0] = pol + vl + en; out[1] = pol; out[2] = vl; out[3] = en;
  }
  cl.sync();  // nobody leaves while its shared memory may still be read
}

inline int launch(const float* logits, const float* baseline, const float* bootstrap, const float* raw_action,
                  const float* behaviour_lp, const float* reward, const float* discount, const float* truncation,
                  const float* noise, int B, int T, int A, Hyper h, float* ws, float* out, float* grad_logits,
                  float* grad_baseline, cudaStream_t stream) {
  if (A > MAXA || A <= 0 || B <= 0 || T <= 0) return 1;
  // cluster version whenever a CTA's slice fits its shared memory (RSRX_PPO_HEAD_CLUSTER=0: the single-CTA kernel)
  static const bool want_cluster = [] { const char* e = getenv("RSRX_PPO_HEAD_CLUSTER"); return !(e && e[0] == '0'); }();
  const int cap = ((B + CL - 1) / CL) * T;
  const size_t smem = sizeof(float) * 4 * (size_t)cap;
  if (want_cluster && smem <= 160 * 1024) {
    static bool set = false;
    if (!set) {
      if (cudaFuncSetAttribute(head_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024) != cudaSuccess) return 1;
      set = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(CL);
    cfg.blockDim = dim3(CT);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CL; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[1].val.programmaticStreamSerializationAllowed = pdl::enabled() ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 2;
    return cudaLaunchKernelEx(&cfg, head_cluster_kernel, logits, baseline, bootstrap, raw_action, behaviour_lp, reward, discount,
                              truncation, noise, B, T, A, h, cap, out, grad_logits, grad_baseline) != cudaSuccess;
  }
  return pdl::launch(head_kernel, dim3(1), dim3(THREADS), 0, stream, logits, baseline, bootstrap, raw_action, behaviour_lp, reward,
                     discount, truncation, noise, B, T, A, h, ws, out, grad_logits, grad_baseline) != cudaSuccess;
}

}  // namespace ppo
}  // namespace rsrx

// ---- MLP backward helper: activation derivative + bias gradient in one launch ---------------------------------------
// grad_z = grad_y * act'(z) (written when grad_z != nullptr) and db[c] = sum_rows grad_z[r][c], deterministically: every
// block writes its partial column sums, the last block to finish adds them in block order.  Replaces torch's
// silu_backward + a dim-0 reduce kernel per layer (10 us -> 3 us on [2816, 256]).
namespace rsrx {
namespace mlp {

constexpr int TX = 32, TY = 8, ROWS_PER_BLOCK = 128;

__device__ __forceinline__ float act_grad(int act, float z) {
  if (act == 1) {  // silu: s (1 + z (1 - s))
    const float s = 1.f / (1.f + expf(-z));
    return s * (1.f + z * (1.f - s));
  }
  if (act == 2) return z > 0.f ? 1.f : 0.f;  // relu
  return 1.f;
}

__global__ void __launch_bounds__(TX * TY) act_bias_backward_kernel(const float* __restrict__ grad_y, const float* __restrict__ z,
                                                                    int rows, int cols, int act, float* __restrict__ grad_z,
                                                                    float* __restrict__ db, float* __restrict__ partial,
                                                                    unsigned* __restrict__ counter) {
  __shared__ float red[TY][TX];
  __shared__ bool last;
  const int c = blockIdx.x * TX + threadIdx.x;
  const int r0 = blockIdx.y * ROWS_PER_BLOCK;
  float s = 0.f;
  if (c < cols) {
    for (int r = r0 + threadIdx.y; r < min(r0 + ROWS_PER_BLOCK, rows); r += TY) {
      const size_t i = (size_t)r * cols + c;
      const float g = grad_y[i] * act_grad(act, act ? z[i] : 0.f);
      if (grad_z) grad_z[i] = g;
      s += g;
    }
  }
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.f;
    for (int y = 0; y < TY; y++) t += red[y][threadIdx.x];
    partial[(size_t)blockIdx.y * cols + c] = t;
  }
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0 && threadIdx.y == 0) last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!last) return;
  __threadfence();
  for (int cc = threadIdx.y * TX + threadIdx.x; cc < cols; cc += TX * TY) {
    float t = 0.f;
    for (unsigned b = 0; b < gridDim.y; b++) t += partial[(size_t)b * cols + cc];
    db[cc] = t;
  }
  if (threadIdx.x == 0 && threadIdx.y == 0) *counter = 0u;
}

inline size_t workspace_floats(int rows, int cols) {
  return (size_t)((rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK) * cols + 1;
}

inline int launch(const float* grad_y, const float* z, int rows, int cols, int act, float* grad_z, float* db,
                  float* workspace, cudaStream_t stream) {
  const dim3 grid((cols + TX - 1) / TX, (rows + ROWS_PER_BLOCK - 1) / ROWS_PER_BLOCK), block(TX, TY);
  unsigned* counter = reinterpret_cast<unsigned*>(workspace);
  act_bias_backward_kernel<<<grid, block, 0, stream>>>(grad_y, z, rows, cols, act, grad_z, db, workspace + 1, counter);
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace mlp
}  // namespace rsrx

// ---- minibatch gather: dst_f[r][:] = src_f[idx[r]][:] for up to 8 row-major float tensors in one launch ---------------
namespace rsrx {
namespace gather {

constexpr int MAXF = 8;
struct Fields {
  const float* src[MAXF];
  float* dst[MAXF];
  int width[MAXF];  // floats per row
  int n;
};

__global__ void __launch_bounds__(256) gather_kernel(Fields f, const long long* __restrict__ idx, int nrows) {
  pdl::launch_dependents();
  pdl::wait();
  const int r = blockIdx.x;
  if (r >= nrows) return;
  const long long s = idx[r];
  for (int k = 0; k < f.n; k++) {
    const float* src = f.src[k] + (size_t)s * f.width[k];
    float* dst = f.dst[k] + (size_t)r * f.width[k];
    for (int i = threadIdx.x; i < f.width[k]; i += blockDim.x) dst[i] = src[i];
  }
}

inline int launch(const Fields& f, const long long* idx, int nrows, cudaStream_t stream) {
  return pdl::launch(gather_kernel, dim3(nrows), dim3(256), 0, stream, f, idx, nrows) != cudaSuccess;
}

}  // namespace gather
}  // namespace rsrx

// ---- behaviour policy head: raw = loc + scale * noise, action = tanh(raw), log-prob of raw, one launch ------------------
// (brax NormalTanhDistribution.sample_no_postprocessing / log_prob / postprocess; RSR/train.py:313 actor_step)
namespace rsrx {
namespace ppo {

__global__ void __launch_bounds__(256) act_kernel(const float* __restrict__ logits, const float* __restrict__ noise, int N, int A,
                                                  float* __restrict__ raw, float* __restrict__ action,
                                                  float* __restrict__ log_prob) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  if (n >= N) return;
  const float* lg = logits + (size_t)n * 2 * A;
  float lp = 0.f;
  for (int k = 0; k < A; k++) {
    const float loc = lg[k], sc = softplus(lg[A + k]) + MIN_STD;
    const float x = loc + sc * noise[(size_t)n * A + k];
    const float z = (x - loc) / sc;
    lp += -0.5f * z * z - 0.5f * LOG_2PI - logf(sc) - log_det_jac(x);
    raw[(size_t)n * A + k] = x;
    action[(size_t)n * A + k] = tanhf(x);
  }
  if (log_prob) log_prob[n] = lp;
}

inline int launch_act(const float* logits, const float* noise, int N, int A, float* raw, float* action, float* log_prob,
                      cudaStream_t stream) {
  act_kernel<<<(N + 255) / 256, 256, 0, stream>>>(logits, noise, N, A, raw, action, log_prob);
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace ppo
}  // namespace rsrx

// ---- minibatch input preparation for the fused PPO update: one launch instead of seven ------------------------------
// From the gathered minibatch (obs / next_obs [mb][T][O]) and the running-statistics normaliser (mean, std [O]):
//   obs_n [mb * T][O]                 normalised observations (policy input)
//   x_pad [mb * T + mb][ldp]          value-network input, zero-padded to ldp columns: the mb * T normalised observations
//                                     followed by the mb normalised bootstrap observations next_obs[:, T - 1]
//   xT    [ldp][ldt]                  its transpose (the contraction-contiguous operand of the first layer's weight gradient)
// 32 x 32 tiles through shared memory so that both layouts are written with contiguous rows.
namespace rsrx {
namespace ppo {

__global__ void __launch_bounds__(256) prep_kernel(const float* __restrict__ obs, const float* __restrict__ next_obs,
                                                   const float* __restrict__ mean, const float* __restrict__ stdv, int mb, int T,
                                                   int O, float* __restrict__ obs_n, float* __restrict__ x_pad, int ldp,
                                                   float* __restrict__ xT, int ldt) {
  __shared__ float tile[32][33];
  pdl::launch_dependents();
  pdl::wait();
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int rows = mb * T + mb, r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const int col = c0 + tx;
  const float mu = col < O ? mean[col] : 0.f, isd = col < O ? 1.f / stdv[col] : 0.f;
  for (int y = ty; y < 32; y += 8) {
    const int row = r0 + y;
    float v = 0.f;
    if (row < rows && col < O) {
      const float* src = row < mb * T ? obs + (size_t)row * O : next_obs + ((size_t)(row - mb * T) * T + (T - 1)) * O;
      v = (src[col] - mu) * isd;
      if (row < mb * T) obs_n[(size_t)row * O + col] = v;
    }
    if (row < rows && col < ldp) x_pad[(size_t)row * ldp + col] = v;
    tile[y][tx] = v;
  }
  __syncthreads();
  for (int y = ty; y < 32; y += 8)
    if (c0 + y < ldp && r0 + tx < rows) xT[(size_t)(c0 + y) * ldt + r0 + tx] = tile[tx][y];
}

inline int launch_prep(const float* obs, const float* next_obs, const float* mean, const float* stdv, int mb, int T, int O,
                       float* obs_n, float* x_pad, int ldp, float* xT, int ldt, cudaStream_t stream) {
  const int rows = mb * T + mb;
  return pdl::launch(prep_kernel, dim3((rows + 31) / 32, (ldp + 31) / 32), dim3(256), 0, stream, obs, next_obs, mean, stdv, mb, T, O, obs_n,
                     x_pad, ldp, xT, ldt) != cudaSuccess;
}

}  // namespace ppo
}  // namespace rsrx
