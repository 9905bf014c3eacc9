// rsrx_loss.cuh — fused RSR distribution loss (forward + backward), one launch.
//   reference: RSR/dataset_processor.py:17-43 (evaluate_kde, wasserstein_distance),
//              RSR/rsr_loss.py:122-175 (compute_rsr_loss)
// density[m]  = softmax_m( logsumexp_n( -|g_m - x_n|^2 / (2 h^2) ) - log N )
// distance    = sum_m | cumsum(density)_m - cumsum(reference_density)_m |
// loss        = loss_scale * divergence * distance
// d loss/d x_n = loss_scale * divergence * sum_m a_m w_mn (g_m - x_n) / h^2,
//   w_mn = exp(logk_mn - lse_m),  a_m = p_m (s_m - sum_j p_j s_j),
//   s_j = sum_{m >= j} sign(cumsum(p)_m - cumsum(q)_m).
// The problem is tiny (M ~ 10 grid points, N ~ 10^3 rows, D = 51 | 108) and
// latency-bound: ONE CTA, one warp per data row (lanes over D: coalesced row
// loads), grid points staged in shared memory, fp32 difference form (a TF32/BF16
// Gram expansion would lose the softmax: logits are -50 * |g - x|^2 ~ 1e3..1e4).
#pragma once
#include <cuda_runtime.h>
#include <math.h>

namespace rsrx {
namespace loss {

constexpr int MAXM = 64;
constexpr int MAXD = 256;
constexpr int THREADS = 1024;
constexpr int NWARP = THREADS / 32;
constexpr int DREG = MAXD / 32;  // per-lane slice of a row

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__device__ __forceinline__ const float* row_ptr(const float* ref, int Nref, const float* batch, int D, int n) {
  return n < Nref ? ref + (size_t)n * D : batch + (size_t)(n - Nref) * D;
}

__global__ void __launch_bounds__(THREADS) rsr_loss_kernel(const float* __restrict__ grid, int M, int D,
                                                          const float* __restrict__ ref, int Nref,
                                                          const float* __restrict__ batch, int Nb,
                                                          const float* __restrict__ refdens, float bandwidth,
                                                          float divergence, float loss_scale,
                                                          float* __restrict__ density_out, float* __restrict__ out,
                                                          float* __restrict__ grad) {
  extern __shared__ float sh[];
  float* sg = sh;                        // [M][D]
  float* wmax = sg + M * D;              // [NWARP][M]
  float* wsumexp = wmax + NWARP * M;     // [NWARP][M]
  float* lse = wsumexp + NWARP * M;      // [M]
  float* am = lse + M;                   // [M]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int N = Nref + Nb;
  const float inv2h2 = 1.f / (2.f * bandwidth * bandwidth);
  for (int i = tid; i < M * D; i += THREADS) sg[i] = grid[i];
  __syncthreads();
  // ---- pass 1: per-warp online logsumexp over its rows
  float rmax[MAXM / 32 + 1], rsum[MAXM / 32 + 1];  // lane l keeps grid points m = l, l+32
#pragma unroll
  for (int k = 0; k < MAXM / 32 + 1; k++) { rmax[k] = -INFINITY; rsum[k] = 0.f; }
  for (int n = warp; n < N; n += NWARP) {
    const float* x = row_ptr(ref, Nref, batch, D, n);
    float xr[DREG];
#pragma unroll
    for (int k = 0; k < DREG; k++) { const int d = lane + 32 * k; xr[k] = d < D ? x[d] : 0.f; }
    for (int m = 0; m < M; m++) {
      float s = 0.f;
#pragma unroll
      for (int k = 0; k < DREG; k++) {
        const int d = lane + 32 * k;
        if (d < D) { const float df = sg[m * D + d] - xr[k]; s += df * df; }
      }
      const float lk = -wsum(s) * inv2h2;
      if ((m & 31) == lane) {
        const int k = m >> 5;
        const float mx = fmaxf(rmax[k], lk);
        rsum[k] = rsum[k] * __expf(rmax[k] - mx) + __expf(lk - mx);
        rmax[k] = mx;
      }
    }
  }
#pragma unroll
  for (int k = 0; k < MAXM / 32 + 1; k++) {
    const int m = lane + 32 * k;
    if (m < M) { wmax[warp * M + m] = rmax[k]; wsumexp[warp * M + m] = rsum[k]; }
  }
  __syncthreads();
  if (tid < M) {
    float mx = -INFINITY;
    for (int w = 0; w < NWARP; w++) mx = fmaxf(mx, wmax[w * M + tid]);
    float s = 0.f;
    for (int w = 0; w < NWARP; w++) {
      const float wm = wmax[w * M + tid];
      if (wm > -INFINITY) s += wsumexp[w * M + tid] * expf(wm - mx);
    }
    lse[tid] = mx + logf(s);  // logsumexp_n; the "- log N" cancels in the softmax
  }
  __syncthreads();
  // ---- density, distance, a_m (one thread; M <= 64)
  if (tid == 0) {
    float mx = -INFINITY;
    for (int m = 0; m < M; m++) mx = fmaxf(mx, lse[m]);
    float z = 0.f;
    float p[MAXM];
    for (int m = 0; m < M; m++) { p[m] = expf(lse[m] - mx); z += p[m]; }
    for (int m = 0; m < M; m++) { p[m] /= z; if (density_out) density_out[m] = p[m]; }
    if (refdens) {
      float cp = 0.f, cq = 0.f, dist = 0.f;
      float sg_[MAXM];
      for (int m = 0; m < M; m++) {
        cp += p[m]; cq += refdens[m];
        const float df = cp - cq;
        dist += fabsf(df);
        sg_[m] = df > 0.f ? 1.f : (df < 0.f ? -1.f : 0.f);
      }
      float suf = 0.f, ps = 0.f;
      for (int m = M - 1; m >= 0; m--) { suf += sg_[m]; sg_[m] = suf; }
      for (int m = 0; m < M; m++) ps += p[m] * sg_[m];
      for (int m = 0; m < M; m++) am[m] = p[m] * (sg_[m] - ps);
      if (out) { out[0] = loss_scale * divergence * dist; out[1] = dist; }
    }
  }
  __syncthreads();
  // ---- pass 2: gradient w.r.t. the online batch rows
  if (grad && refdens) {
    const float coef = loss_scale * divergence / (bandwidth * bandwidth);
    for (int n = Nref + warp; n < N; n += NWARP) {
      const float* x = batch + (size_t)(n - Nref) * D;
      float xr[DREG], gr[DREG];
#pragma unroll
      for (int k = 0; k < DREG; k++) { const int d = lane + 32 * k; xr[k] = d < D ? x[d] : 0.f; gr[k] = 0.f; }
      for (int m = 0; m < M; m++) {
        float s = 0.f;
#pragma unroll
        for (int k = 0; k < DREG; k++) {
          const int d = lane + 32 * k;
          if (d < D) { const float df = sg[m * D + d] - xr[k]; s += df * df; }
        }
        const float lk = -wsum(s) * inv2h2;
        const float c = am[m] * expf(lk - lse[m]);
#pragma unroll
        for (int k = 0; k < DREG; k++) {
          const int d = lane + 32 * k;
          if (d < D) gr[k] += c * (sg[m * D + d] - xr[k]);
        }
      }
#pragma unroll
      for (int k = 0; k < DREG; k++) {
        const int d = lane + 32 * k;
        if (d < D) grad[(size_t)(n - Nref) * D + d] = coef * gr[k];
      }
    }
  }
}

// returns non-zero on launch failure
inline int launch(const float* grid, int M, int D, const float* ref, int Nref, const float* batch, int Nb,
                  const float* refdens, float bandwidth, float divergence, float loss_scale, float* density_out,
                  float* out, float* grad, cudaStream_t stream) {
  if (D > MAXD) return 1;
  const size_t smem = sizeof(float) * ((size_t)M * D + 2 * NWARP * M + 2 * M);
  static bool attr_set = false;
  if (!attr_set) {
    cudaFuncSetAttribute(rsr_loss_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    attr_set = true;
  }
  rsr_loss_kernel<<<1, THREADS, smem, stream>>>(grid, M, D, ref, Nref, batch, Nb, refdens, bandwidth, divergence,
                                               loss_scale, density_out, out, grad);
  return cudaGetLastError() != cudaSuccess;
}

}  // namespace loss
}  // namespace rsrx
