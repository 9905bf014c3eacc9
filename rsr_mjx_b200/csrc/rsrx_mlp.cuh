// rsrx_mlp.cuh — the trainers' POLICY network (brax make_ppo_networks policy_hidden_layer_sizes (32,)*4: obs -> 32 x 4 ->
// 2 * action_size, swish; RSR/train.py) forward and backward as one launch each.
//
// The network is far too thin for the tensor core (widths <= 32) and through torch it is ~45 launches per minibatch step
// (an addmm + SiLU per layer forward; SiLU' + bias reduction + two tiny cuBLAS GEMMs per layer backward: 0.2 ms).
// Here: one WARP per batch row, LANE = neuron.  Activations live in one register per lane, an input h[k] reaches every
// lane by shuffle, weights sit in shared memory laid out so that lanes read consecutive words.
//   forward : z[j] = b[j] + sum_k h[k] W[j][k]; pre-activations of the hidden layers are kept for the backward pass
//   backward: dz -> weight / bias gradients accumulated per warp in shared memory (lane j owns column j: no atomics),
//             dh[k] = sum_j dz[j] W[j][k], dz_prev = dh * act'(z_prev); the CTA adds its warps in a fixed order and writes
//             one partial gradient vector per CTA, finished by rsrx_reduce_partials (deterministic).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "rsrx_pdl.cuh"

namespace rsrx {
namespace smallmlp {

constexpr int MAXL = 8;     // layers
constexpr int WD = 32;      // widest layer (= lanes)
constexpr int WARPS = 8;

struct Net {
  const float* W[MAXL];     // [out][in] row-major (torch nn.Linear)
  const float* b[MAXL];     // [out]
  int width[MAXL + 1];      // width[0] = inputs, width[l + 1] = outputs of layer l
  int nl, act;              // act: 1 silu, 2 relu (hidden layers; the output layer is linear)
};

__device__ __forceinline__ float act_f(int act, float z) { return act == 1 ? z / (1.f + expf(-z)) : (act == 2 ? fmaxf(z, 0.f) : z); }
__device__ __forceinline__ float act_d(int act, float z) {
  if (act == 1) { const float s = 1.f / (1.f + expf(-z)); return s * (1.f + z * (1.f - s)); }
  return act == 2 ? (z > 0.f ? 1.f : 0.f) : 1.f;
}

// zs: [nl - 1][rows][WD] pre-activations of the hidden layers; out: [rows][ldo]
// nmean / nstd (optional, [width[0]]): the input is normalised on the way in, (x - mean) / std — the actor step of the
// trainers (running-statistics normaliser, RSR/train.py:313 through acting.actor_step) without a separate launch
__global__ void __launch_bounds__(32 * WARPS) forward_kernel(const Net net, const float* __restrict__ x, int ldx, int rows,
                                                            float* __restrict__ zs, float* __restrict__ out, int ldo,
                                                            const float* __restrict__ nmean, const float* __restrict__ nstd) {
  extern __shared__ float sm[];  // Wt[l][k][j] = W[l][j][k], padded to WD x WD; then bias[l][WD]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* sb = sm + net.nl * WD * WD;
  pdl::launch_dependents();
  // all weights of all layers in flight at once (4-byte cp.async straight into the transposed, zero-padded image): one
  // global round trip for the prologue instead of one per layer
  for (int i = tid; i < net.nl * WD * WD; i += blockDim.x) sm[i] = 0.f;
  __syncthreads();
  pdl::wait();  // the zero fill above overlaps the predecessor's tail
  for (int l = 0; l < net.nl; ++l) {
    const int win = net.width[l], wout = net.width[l + 1];
    for (int i = tid; i < win * WD; i += blockDim.x) {  // shared-memory order: consecutive lanes write consecutive words
      const int k = i / WD, j = i % WD;
      if (j < wout)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sm + l * WD * WD + i)),
                     "l"(net.W[l] + j * win + k) : "memory");
    }
    if (tid < WD) sb[l * WD + tid] = tid < wout ? net.b[l][tid] : 0.f;
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int row = blockIdx.x * WARPS + warp; row < rows; row += gridDim.x * WARPS) {
    float h = lane < net.width[0] ? x[(size_t)row * ldx + lane] : 0.f;
    if (nmean && lane < net.width[0]) h = (h - nmean[lane]) / nstd[lane];
    for (int l = 0; l < net.nl; ++l) {
      const int win = net.width[l];
      const float* wt = sm + l * WD * WD;
      // four partial sums: the 32-term dot product is otherwise one chain of dependent FMAs behind shared-memory loads
      float z0 = sb[l * WD + lane], z1 = 0.f, z2 = 0.f, z3 = 0.f;
#pragma unroll 2
      for (int k = 0; k < win; k += 4) {  // rows k >= win of wt are zero, h of lanes >= win is zero or unused
        z0 += __shfl_sync(0xffffffffu, h, k) * wt[k * WD + lane];
        z1 += __shfl_sync(0xffffffffu, h, k + 1) * wt[(k + 1) * WD + lane];
        z2 += __shfl_sync(0xffffffffu, h, k + 2) * wt[(k + 2) * WD + lane];
        z3 += __shfl_sync(0xffffffffu, h, k + 3) * wt[(k + 3) * WD + lane];
      }
      const float z = (z0 + z1) + (z2 + z3);
      if (l + 1 < net.nl) {
        zs[((size_t)l * rows + row) * WD + lane] = z;
        h = act_f(net.act, z);
      } else if (lane < net.width[l + 1]) {
        out[(size_t)row * ldo + lane] = z;
      }
    }
  }
}

// g: [rows][ldg] gradient w.r.t. the outputs; partials: [gridDim.x][total] in parameter order (W_0, b_0, W_1, b_1, ...)
__global__ void __launch_bounds__(32 * WARPS) backward_kernel(const Net net, const float* __restrict__ x, int ldx, int rows,
                                                             const float* __restrict__ zs, const float* __restrict__ g, int ldg,
                                                             float* __restrict__ partials, int total) {
  extern __shared__ float sm[];  // Ws[l][j][k] = W[l][j][k] padded; then per warp: accW[l][k][j], accb[l][j]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nl = net.nl, per_warp = nl * WD * WD + nl * WD;
  float* sW = sm;
  float* acc = sm + nl * WD * WD + warp * per_warp;
  pdl::launch_dependents();
  for (int i = tid; i < nl * WD * WD; i += blockDim.x) sW[i] = 0.f;
  for (int i = lane; i < per_warp; i += 32) acc[i] = 0.f;
  __syncthreads();
  pdl::wait();  // the accumulator zero fill (0.2 MB of shared memory per CTA) overlaps the predecessor's tail
  for (int l = 0; l < nl; ++l) {  // every layer's weights in flight at once
    const int win = net.width[l], wout = net.width[l + 1];
    for (int i = tid; i < wout * win; i += blockDim.x)
      asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((uint32_t)__cvta_generic_to_shared(sW + l * WD * WD + (i / win) * WD + i % win)),
                   "l"(net.W[l] + i) : "memory");
  }
  asm volatile("cp.async.wait_all;" ::: "memory");
  __syncthreads();
  for (int row = blockIdx.x * WARPS + warp; row < rows; row += gridDim.x * WARPS) {
    // everything this row reads from global memory is requested up front (one latency per row instead of one per layer)
    float dz = lane < net.width[nl] ? g[(size_t)row * ldg + lane] : 0.f;
    const float xr = lane < net.width[0] ? x[(size_t)row * ldx + lane] : 0.f;
    float zr[MAXL - 1];
#pragma unroll
    for (int l = 0; l < MAXL - 1; ++l) zr[l] = l < nl - 1 ? zs[((size_t)l * rows + row) * WD + lane] : 0.f;
#pragma unroll
    for (int l = MAXL - 1; l >= 0; --l) {
      if (l >= nl) continue;
      const int win = net.width[l], wout = net.width[l + 1];
      float zprev = 0.f, hprev;
      if (l == 0) hprev = xr;
      else { zprev = zr[l - 1]; hprev = act_f(net.act, zprev); }
      float* aW = acc + l * WD * WD;
      // dW[j][k] += dz[j] h[k]: lane j owns column j of the transposed accumulator
      // read-modify-write in batches of 8 independent entries (the compiler cannot prove the accumulators do not alias
      // the weights, so a plain loop serialises load -> FMA -> store)
#pragma unroll 1
      for (int k0 = 0; k0 < win; k0 += 8) {
        float t8[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) t8[u] = aW[(k0 + u) * WD + lane];
#pragma unroll
        for (int u = 0; u < 8; ++u) t8[u] += dz * __shfl_sync(0xffffffffu, hprev, k0 + u);  // lanes >= win hold hprev = 0
#pragma unroll
        for (int u = 0; u < 8; ++u) aW[(k0 + u) * WD + lane] = t8[u];
      }
      acc[nl * WD * WD + l * WD + lane] += dz;
      if (l > 0) {
        const float* w = sW + l * WD * WD;
        float d0 = 0.f, d1 = 0.f, d2 = 0.f, d3 = 0.f;
#pragma unroll 2
        for (int j = 0; j < wout; j += 4) {  // rows j >= wout of w are zero
          d0 += __shfl_sync(0xffffffffu, dz, j) * w[j * WD + lane];
          d1 += __shfl_sync(0xffffffffu, dz, j + 1) * w[(j + 1) * WD + lane];
          d2 += __shfl_sync(0xffffffffu, dz, j + 2) * w[(j + 2) * WD + lane];
          d3 += __shfl_sync(0xffffffffu, dz, j + 3) * w[(j + 3) * WD + lane];
        }
        dz = lane < win ? ((d0 + d1) + (d2 + d3)) * act_d(net.act, zprev) : 0.f;
      }
    }
  }
  __syncthreads();
  // CTA partial in parameter order; warps added in index order
  float* outp = partials + (size_t)blockIdx.x * total;
  int off = 0;
  const float* acc0 = sm + nl * WD * WD;
  for (int l = 0; l < nl; ++l) {
    const int win = net.width[l], wout = net.width[l + 1];
    // walk the accumulators in THEIR order (k-major: consecutive threads read consecutive words; the parameter-order walk
    // was a 32-way bank conflict and 20 us per launch) and scatter the sums to parameter order
    for (int i = tid; i < win * WD; i += blockDim.x) {
      const int k = i / WD, j = i % WD;
      if (j >= wout) continue;
      float t = 0.f;
      for (int w = 0; w < WARPS; ++w) t += acc0[w * per_warp + l * WD * WD + i];
      outp[off + j * win + k] = t;
    }
    off += wout * win;
    for (int j = tid; j < wout; j += blockDim.x) {
      float t = 0.f;
      for (int w = 0; w < WARPS; ++w) t += acc0[w * per_warp + nl * WD * WD + l * WD + j];
      outp[off + j] = t;
    }
    off += wout;
  }
}

inline size_t fwd_smem(int nl) { return sizeof(float) * (size_t)(nl * WD * WD + nl * WD); }
inline size_t bwd_smem(int nl) { return sizeof(float) * (size_t)(nl * WD * WD + WARPS * (nl * WD * WD + nl * WD)); }

}  // namespace smallmlp
}  // namespace rsrx
