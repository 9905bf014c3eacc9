// rsrx_loss.cuh — RSR distribution loss, forward + backward, spread over the chip (two short launches).
//   reference: RSR/dataset_processor.py:17-43 (evaluate_kde, wasserstein_distance),
//              RSR/rsr_loss.py:122-175 (compute_rsr_loss)
// density[m]  = softmax_m( logsumexp_n( -|g_m - x_n|^2 / (2 h^2) ) - log N )
// distance    = sum_m | cumsum(density)_m - cumsum(reference_density)_m |
// loss        = loss_scale * divergence * distance
// d loss/d x_n = loss_scale * divergence * sum_m a_m w_mn (g_m - x_n) / h^2,
//   w_mn = exp(logk_mn - lse_m),  a_m = p_m (s_m - sum_j p_j s_j),
//   s_j = sum_{m >= j} sign(cumsum(p)_m - cumsum(q)_m).
// The problem is small (M ~ 10 grid points, N ~ 10^3 rows, D = 51 | 108) and latency-bound, so it is spread over the
// chip in two short launches:
//   kde_kernel   ~N/32 CTAs x 8 warps, one warp per data row (lanes over D: coalesced row loads; grid points in shared
//                memory), logk[n][m] kept in a workspace, online logsumexp per warp -> per CTA -> the LAST CTA to finish
//                (ticket) combines the CTA partials and does the M-sized tail — softmax, cumsum, |.| sum, sign suffix
//                sums, a_m — on one warp with shuffle scans;
//   grad_kernel  one warp per online-batch row: grad_n = coef * (sum_m c_mn g_m - x_n sum_m c_mn), c_mn = a_m exp(logk_mn
//                - lse_m) from the stored logk: no reductions at all.
// fp32 difference form throughout (a TF32/BF16 Gram expansion would lose the softmax: logits are -50 |g - x|^2 ~
// 1e3..1e4), full-precision expf in both passes.
#pragma once
#include <cuda_runtime.h>
#include <math.h>

#include <algorithm>

#include "rsrx_pdl.cuh"

namespace rsrx {
namespace loss {

constexpr int MAXM = 64;
constexpr int MAXD = 256;
constexpr int WARPS = 8;
constexpr int THREADS = 32 * WARPS;
constexpr int DREG = MAXD / 32;  // per-lane slice of a row
constexpr int MAXCTA = 148;

// device workspace of one call (library-owned, see launch()): logk [N][M] | partials [MAXCTA][M][2] | lse [M] | am [M] | ticket
struct Workspace {
  float* logk;
  float* part;
  float* lse;
  float* am;
  unsigned* ticket;
};

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float wmaxf(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// inclusive scan over the 64 values a warp holds as (lo = element lane, hi = element lane + 32)
__device__ __forceinline__ void scan64(float& lo, float& hi, int lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float a = __shfl_up_sync(0xffffffffu, lo, o), b = __shfl_up_sync(0xffffffffu, hi, o);
    if (lane >= o) { lo += a; hi += b; }
  }
  hi += __shfl_sync(0xffffffffu, lo, 31);
}

__device__ __forceinline__ const float* row_ptr(const float* ref, int Nref, const float* batch, int D, int n) {
  return n < Nref ? ref + (size_t)n * D : batch + (size_t)(n - Nref) * D;
}

__global__ void __launch_bounds__(THREADS) kde_kernel(const float* __restrict__ grid, int M, int D,
                                                     const float* __restrict__ ref, int Nref,
                                                     const float* __restrict__ batch, int Nb,
                                                     const float* __restrict__ refdens, float bandwidth,
                                                     float divergence, float loss_scale,
                                                     float* __restrict__ density_out, float* __restrict__ out, Workspace ws) {
  extern __shared__ float sh[];
  float* sg = sh;                        // [M][D]
  float* wmx = sg + max(M * D, (int)gridDim.x * M * 2);  // [WARPS][M]
  float* wsm = wmx + WARPS * M;          // [WARPS][M]
  __shared__ bool last;
  pdl::launch_dependents();
  pdl::wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = Nref + Nb;
  const float inv2h2 = 1.f / (2.f * bandwidth * bandwidth);
  for (int i = tid; i < M * D; i += THREADS) sg[i] = grid[i];
  __syncthreads();
  // ---- per-warp online logsumexp over its rows; lane l keeps grid points m = l, l + 32
  float rmax[2] = {-INFINITY, -INFINITY}, rsum[2] = {0.f, 0.f};
  for (int n = blockIdx.x * WARPS + warp; n < N; n += gridDim.x * WARPS) {
    const float* x = row_ptr(ref, Nref, batch, D, n);
    float xr[DREG];
#pragma unroll
    for (int k = 0; k < DREG; k++) { const int d = lane + 32 * k; xr[k] = d < D ? x[d] : 0.f; }
    float mine[2] = {0.f, 0.f};
    for (int m = 0; m < M; m++) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DREG; k++) {
        const int d = lane + 32 * k;
        if (d < D) { const float df = sg[m * D + d] - xr[k]; s += df * df; }
      }
      const float lk = -wsum(s) * inv2h2;
      if ((m & 31) == lane) mine[m >> 5] = lk;
    }
#pragma unroll
    for (int k = 0; k < 2; k++) {
      const int m = lane + 32 * k;
      if (m < M) {
        ws.logk[(size_t)n * M + m] = mine[k];
        const float mx = fmaxf(rmax[k], mine[k]);
        rsum[k] = rsum[k] * expf(rmax[k] - mx) + expf(mine[k] - mx);
        rmax[k] = mx;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 2; k++) {
    const int m = lane + 32 * k;
    if (m < M) { wmx[warp * M + m] = rmax[k]; wsm[warp * M + m] = rsum[k]; }
  }
  __syncthreads();
  if (tid < M) {  // CTA partial of grid point tid
    float mx = -INFINITY;
    for (int w = 0; w < WARPS; w++) mx = fmaxf(mx, wmx[w * M + tid]);
    float s = 0.f;
    for (int w = 0; w < WARPS; w++) {
      const float wm = wmx[w * M + tid];
      if (wm > -INFINITY) s += wsm[w * M + tid] * expf(wm - mx);
    }
    ws.part[((size_t)blockIdx.x * M + tid) * 2] = mx;
    ws.part[((size_t)blockIdx.x * M + tid) * 2 + 1] = s;
  }
  __threadfence();
  __syncthreads();
  if (tid == 0) last = atomicAdd(ws.ticket, 1u) == gridDim.x - 1;
  __syncthreads();
  if (!last) return;
  // ---- the last CTA: combine the CTA partials, then the M-sized tail on warp 0.  The partials are first pulled into
  // shared memory by all threads at once (independent L2 loads: a serial walk over the CTAs was 80 % of the kernel).
  __threadfence();
  const int G = (int)gridDim.x;
  float* pp = sg;  // [G][M][2] (the grid points are no longer needed by this CTA); capacity checked by launch()
  for (int i = tid; i < G * M * 2; i += THREADS) pp[i] = __ldcg(ws.part + i);
  __syncthreads();
  float* lse = wmx;  // reuse
  for (int m = warp; m < M; m += WARPS) {  // one warp per grid point, lanes over the CTAs
    float mx = -INFINITY;
    for (int c = lane; c < G; c += 32) mx = fmaxf(mx, pp[(c * M + m) * 2]);
    mx = wmaxf(mx);
    float sacc = 0.f;
    for (int c = lane; c < G; c += 32) {
      const float pm = pp[(c * M + m) * 2];
      if (pm > -INFINITY) sacc += pp[(c * M + m) * 2 + 1] * expf(pm - mx);
    }
    sacc = wsum(sacc);
    if (lane == 0) {
      const float l = mx + logf(sacc);  // logsumexp_n; the "- log N" cancels in the softmax
      lse[m] = l;
      ws.lse[m] = l;
    }
  }
  __syncthreads();
  if (warp == 0) {
    const bool h0 = lane < M, h1 = lane + 32 < M;
    const float l0 = h0 ? lse[lane] : -INFINITY, l1 = h1 ? lse[lane + 32] : -INFINITY;
    const float mx = wmaxf(fmaxf(l0, l1));
    float p0 = h0 ? expf(l0 - mx) : 0.f, p1 = h1 ? expf(l1 - mx) : 0.f;
    const float z = wsum(p0 + p1);
    p0 /= z; p1 /= z;
    if (density_out) { if (h0) density_out[lane] = p0; if (h1) density_out[lane + 32] = p1; }
    if (refdens) {
      // cumsum(p) - cumsum(q) in the reference's form: two scans, then the difference
      float cp0 = p0, cp1 = p1, cq0 = h0 ? refdens[lane] : 0.f, cq1 = h1 ? refdens[lane + 32] : 0.f;
      scan64(cp0, cp1, lane);
      scan64(cq0, cq1, lane);
      const float c0 = cp0 - cq0, c1 = cp1 - cq1;
      const float dist = wsum((h0 ? fabsf(c0) : 0.f) + (h1 ? fabsf(c1) : 0.f));
      const float g0 = h0 ? (c0 > 0.f ? 1.f : (c0 < 0.f ? -1.f : 0.f)) : 0.f;
      const float g1 = h1 ? (c1 > 0.f ? 1.f : (c1 < 0.f ? -1.f : 0.f)) : 0.f;
      // s_j = sum_{m >= j} sign_m = total - (inclusive prefix - own)
      float i0 = g0, i1 = g1;
      scan64(i0, i1, lane);
      const float tot = __shfl_sync(0xffffffffu, i1, 31);
      const float s0 = tot - (i0 - g0), s1 = tot - (i1 - g1);
      const float ps = wsum(p0 * s0 + p1 * s1);
      if (h0) ws.am[lane] = p0 * (s0 - ps);
      if (h1) ws.am[lane + 32] = p1 * (s1 - ps);
      if (lane == 0 && out) { out[0] = loss_scale * divergence * dist; out[1] = dist; }
    }
    if (lane == 0) *ws.ticket = 0u;
  }
}

// gradient w.r.t. the online batch rows
__global__ void __launch_bounds__(THREADS) grad_kernel(const float* __restrict__ grid, int M, int D, int Nref,
                                                      const float* __restrict__ batch, int Nb, float bandwidth,
                                                      float divergence, float loss_scale, float* __restrict__ grad,
                                                      Workspace ws) {
  extern __shared__ float sh[];
  float* sg = sh;             // [M][D]
  float* lse = sg + M * D;    // [M]
  float* am = lse + M;        // [M]
  pdl::launch_dependents();
  pdl::wait();
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  for (int i = tid; i < M * D; i += THREADS) sg[i] = grid[i];
  if (tid < M) { lse[tid] = ws.lse[tid]; am[tid] = ws.am[tid]; }
  __syncthreads();
  const float coef = loss_scale * divergence / (bandwidth * bandwidth);
  for (int r = blockIdx.x * WARPS + warp; r < Nb; r += gridDim.x * WARPS) {
    const float* x = batch + (size_t)r * D;
    const float* lk = ws.logk + (size_t)(Nref + r) * M;
    float gr[DREG], csum = 0.f;
#pragma unroll
    for (int k = 0; k < DREG; k++) gr[k] = 0.f;
    for (int m = 0; m < M; m++) {
      const float c = am[m] * expf(lk[m] - lse[m]);
      csum += c;
#pragma unroll
      for (int k = 0; k < DREG; k++) {
        const int d = lane + 32 * k;
        if (d < D) gr[k] += c * sg[m * D + d];
      }
    }
#pragma unroll
    for (int k = 0; k < DREG; k++) {
      const int d = lane + 32 * k;
      if (d < D) grad[(size_t)r * D + d] = coef * (gr[k] - csum * x[d]);
    }
  }
}

// Library-owned workspace, one per device, grown on demand (an allocation only the first time a larger N * M is seen:
// call once before capturing a CUDA graph).  A block that has been handed out is never freed — a captured graph may
// still point into it — so growth leaves the old block behind (geometric growth: at most ~3x the final size in total).
// Calls on one device must be stream-ordered with each other.
struct WsPool {
  float* base = nullptr;
  size_t floats = 0;
};
inline Workspace workspace(int N, int M, cudaError_t* err) {
  static WsPool pool[64];
  Workspace w{};
  int dev = 0;
  *err = cudaGetDevice(&dev);
  if (*err != cudaSuccess || dev < 0 || dev >= 64) { if (*err == cudaSuccess) *err = cudaErrorInvalidDevice; return w; }
  const size_t fixed = (size_t)MAXCTA * MAXM * 2 + 2 * (size_t)MAXM + 4;
  const size_t need = fixed + (size_t)N * M;
  WsPool& p = pool[dev];
  if (need > p.floats) {
    float* fresh = nullptr;
    const size_t cap = need + need / 2;
    *err = cudaMalloc(&fresh, cap * sizeof(float));
    if (*err != cudaSuccess) return w;
    *err = cudaMemset(fresh, 0, cap * sizeof(float));
    if (*err != cudaSuccess) return w;
    p.base = fresh;
    p.floats = cap;
  }
  w.part = p.base;
  w.lse = w.part + (size_t)MAXCTA * MAXM * 2;
  w.am = w.lse + MAXM;
  w.ticket = reinterpret_cast<unsigned*>(w.am + MAXM);
  w.logk = w.am + MAXM + 4;
  return w;
}

// returns non-zero on failure (cudaGetLastError holds the reason)
inline int launch(const float* grid, int M, int D, const float* ref, int Nref, const float* batch, int Nb,
                  const float* refdens, float bandwidth, float divergence, float loss_scale, float* density_out,
                  float* out, float* grad, cudaStream_t stream) {
  if (D > MAXD || M > MAXM) return 1;
  const int N = Nref + Nb;
  cudaError_t e;
  const Workspace ws = workspace(N, M, &e);
  if (e != cudaSuccess) return 1;
  int g1 = (N + 4 * WARPS - 1) / (4 * WARPS);
  g1 = g1 < 1 ? 1 : (g1 > MAXCTA ? MAXCTA : g1);
  const size_t sg_floats = std::max((size_t)M * D, (size_t)g1 * M * 2);  // grid points, later the CTA partials
  const size_t smem1 = sizeof(float) * (sg_floats + 2 * WARPS * M);
  // only the largest problems (M * D > 12 K floats) need more than the default 48 KB of dynamic shared memory; asking
  // for it unconditionally changes the SM's shared-memory carve-out between neighbouring kernels for nothing
  static bool attr_set = false;
  if (!attr_set && smem1 > 48 * 1024) {
    cudaFuncSetAttribute(kde_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    cudaFuncSetAttribute(grad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  if (pdl::launch(kde_kernel, dim3(g1), dim3(THREADS), smem1, stream, grid, M, D, ref, Nref, batch, Nb, refdens, bandwidth, divergence,
                  loss_scale, density_out, out, ws) != cudaSuccess)
    return 1;
  if (grad && refdens) {
    int g2 = (Nb + 2 * WARPS - 1) / (2 * WARPS);
    g2 = g2 < 1 ? 1 : (g2 > 2 * MAXCTA ? 2 * MAXCTA : g2);
    const size_t smem2 = sizeof(float) * ((size_t)M * D + 2 * M);
    if (pdl::launch(grad_kernel, dim3(g2), dim3(THREADS), smem2, stream, grid, M, D, Nref, batch, Nb, bandwidth, divergence, loss_scale,
                    grad, ws) != cudaSuccess)
      return 1;
  }
  return 0;
}

// ---- the policy-side glue of the RSR term (RSR/losses.py:186-195) without torch in between --------------------------------
// pack:   transition[r] = [obs[r] | tanh(logits[r][0:A]) | next_obs[r]]   (the loss acts on the MODE of the tanh-normal policy)
// unpack: g_out[r][k] = g_in[r][k] + (k < A ? g_transition[r][O + k] * (1 - tanh^2) : 0)   (chain rule through the mode; the
//         scale half of the logits gets no RSR gradient)
__global__ void __launch_bounds__(256) pack_kernel(const float* __restrict__ obs, const float* __restrict__ logits,
                                                   const float* __restrict__ next_obs, int rows, int O, int A,
                                                   float* __restrict__ transition) {
  pdl::launch_dependents();
  pdl::wait();
  const int D = 2 * O + A;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * D; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), c = (int)(i % D);
    float v;
    if (c < O) v = obs[(size_t)r * O + c];
    else if (c < O + A) v = tanhf(logits[(size_t)r * 2 * A + (c - O)]);
    else v = next_obs[(size_t)r * O + (c - O - A)];
    transition[i] = v;
  }
}
__global__ void __launch_bounds__(256) unpack_kernel(const float* __restrict__ transition, const float* __restrict__ g_transition,
                                                     const float* __restrict__ g_in, int rows, int O, int A,
                                                     float* __restrict__ g_out) {
  pdl::launch_dependents();
  pdl::wait();
  const int D = 2 * O + A;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < (size_t)rows * 2 * A; i += (size_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / (2 * A)), k = (int)(i % (2 * A));
    float g = g_in ? g_in[i] : 0.f;
    if (k < A) {
      const float t = transition[(size_t)r * D + O + k];
      g += g_transition[(size_t)r * D + O + k] * (1.f - t * t);
    }
    g_out[i] = g;
  }
}

}  // namespace loss
}  // namespace rsrx
