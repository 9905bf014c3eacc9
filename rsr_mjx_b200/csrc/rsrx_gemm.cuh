// rsrx_gemm.cuh — tcgen05 (5th-gen tensor core) TF32 GEMM with fused epilogues for the PPO / SAC value networks.
//
//   reference: the value network of RSR/train.py (brax make_ppo_networks, 23 -> 256 x 5 -> 1 on 2816 rows per minibatch
//   step, RSR/losses.py:128-131) — the one real contraction on the training path.  Until round 2 it went through
//   torch.addmm + separate SiLU / SiLU' / bias-gradient kernels (40 launches and 0.45 ms per minibatch step).
//
// One kernel, three uses (D = A · B^T accumulated in fp32 in TMEM; inputs are fp32 words read as TF32):
//   forward   Z = X W^T + b,  Y = act(Z)                  A = X [M, K]                B = W [N, K]
//   dgrad     dX = dZ W,      dZprev = dX * act'(Zprev)   A = dZ [M, N']              B[k'][n'] = W[n'][k'] (transposed load)
//             + per-CTA column sums of dZprev (bias gradient partials)
//   wgrad     dW[s] = dZ_s^T X_s over a slice of the rows  A[n'][m] = dZ[m][n'], B[k'][m] = X[m][k'] (both transposed loads),
//             split-K over blockIdx.z, partial tiles to a workspace (summed in order by reduce_partials_kernel:
//             deterministic, no atomics)
//
// Structure (deliberately plain: these GEMMs are 0.4 GFLOP, 44-88 CTAs, latency-bound):
//   CTA = 128 threads, tile 128 x BN x 32, accumulators in TMEM (BN columns), 2 shared-memory stages.
//   All threads stage the operand tiles global -> registers -> shared memory in the canonical no-swizzle UMMA layout
//   (8 x 16-byte core matrices, K-major: ((8,n),2):((16B,SBO),LBO); operands whose global layout is contiguous along
//   M/N instead of K are transposed in flight),
//   fence.proxy.async, one elected thread issues 4 x tcgen05.mma.kind::tf32 (K = 8 each) and tcgen05.commit's them to the
//   stage's mbarrier; the next fill of that stage waits on it.  Epilogue: tcgen05.ld 32x32b (thread = accumulator row),
//   fused bias / activation / activation-derivative / column sums, row-major stores.
//   No TMA: the operands are small, L2-resident activations whose rows are not all 16-byte multiples apart; plain
//   coalesced 16-byte loads keep the kernel free of tensor-map plumbing.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace rsrx {
namespace gemm {

constexpr int BM = 128, BK = 32, THREADS = 128, STAGES = 2;
constexpr int UMMA_K = 8;  // tf32

enum Epilogue { EPI_BIAS_ACT = 0, EPI_DGRAD = 1, EPI_PARTIAL = 2 };
enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2 };

struct Params {
  const float* A; int a_row, a_col;   // element (i, k) of A at A[i * a_row + k * a_col]; one of the two strides is 1
  const float* B; int b_row, b_col;   // element (j, k) of B at B[j * b_row + k * b_col]
  int M, N, K;                        // D is M x N, contraction length K (per split)
  int k_split;                        // wgrad: contraction rows per blockIdx.z slice (multiple of BK); else K
  int epilogue, act;
  const float* bias;                  // [N]            (EPI_BIAS_ACT)
  const float* zprev;                 // [M][ldd]       (EPI_DGRAD: pre-activation of the previous layer)
  float* D;                           // [M][ldd]       Y | dZprev | partial tiles [splits][M][ldd]
  float* Z;                           // [M][ldd]       (EPI_BIAS_ACT: pre-activation, may be null)
  float* colsum;                      // [gridDim.x][ldd] (EPI_DGRAD: per-CTA column sums of dZprev, may be null)
  int ldd;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
// bounded wait: a protocol error traps instead of hanging the GPU
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  for (int spin = 0; spin < (1 << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    if (done) return;
  }
  __trap();
}

__device__ __forceinline__ float act_fwd(int act, float z) {
  if (act == ACT_SILU) return z / (1.f + expf(-z));
  if (act == ACT_RELU) return z > 0.f ? z : 0.f;
  return z;
}
__device__ __forceinline__ float act_bwd(int act, float z) {
  if (act == ACT_SILU) { const float s = 1.f / (1.f + expf(-z)); return s * (1.f + z * (1.f - s)); }
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

// shared-memory (matrix) descriptor, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

// Stage one operand tile (ROWS x BK, row index = the operand's M/N index, column = contraction index) into shared memory.
// The shared-memory image is always K-major: core matrix = 8 rows x 16 B; element (r, k) at byte
//       (r % 8) * 16 + (r / 8) * SBO + (k / 4) * 128 + (k % 4) * 4,  SBO = BK / 4 * 128, LBO = 128.
//   MN_MAJOR = false: global memory contiguous along the contraction: 16-byte loads land as they are;
//   MN_MAJOR = true: global memory contiguous along the row index: transposed in flight (see below).
// Rows >= rows_valid and contraction indices >= k_valid are zero-filled.  16-byte global loads: the contiguous stride
// is 1 and every row / k start is 16-byte aligned (host-checked).
template <int ROWS, bool MN_MAJOR>
__device__ __forceinline__ void stage_tile(float* smem, const float* __restrict__ g, int row_stride, int col_stride, int row0,
                                           int k0, int rows_valid, int k_valid) {
  const int tid = threadIdx.x;
  if (!MN_MAJOR) {
    // thread -> (row, 16-byte chunk of the contraction): 8 consecutive threads take the 8 rows of a core matrix
    constexpr int CH = BK / 4;  // chunks per row
#pragma unroll
    for (int it = 0; it < ROWS * CH / THREADS; ++it) {
      const int idx = it * THREADS + tid;
      const int r = idx % ROWS, c = idx / ROWS;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + r < rows_valid && k0 + c * 4 < k_valid)
        v = *reinterpret_cast<const float4*>(g + (size_t)(row0 + r) * row_stride + (k0 + c * 4));
      *reinterpret_cast<float4*>(smem + ((r & 7) * 4 + (r >> 3) * (CH * 32) + c * 32)) = v;
    }
  } else {
    // global memory is contiguous along the row index: 16-byte loads of 4 rows at one contraction index, transposed on
    // the way into the same K-major image (4 scalar stores).  Lane bits: [0:1] k & 3, [2] chunk parity, [3:4] (k >> 2) & 3
    // -> a warp reads 16 contraction rows x 32 contiguous bytes; the 4 stores are issued in an order rotated by
    // (k >> 2) & 3 so that the 32 lanes of one store instruction hit 32 distinct banks.
    constexpr int CH = BK / 4, RC = ROWS / 4;
#pragma unroll
    for (int it = 0; it < RC * BK / THREADS; ++it) {
      const int idx = it * THREADS + tid;
      const int hi = idx >> 5;
      const int k = (idx & 3) | (((idx >> 3) & 3) << 2) | ((hi & 1) << 4);
      const int rc = ((idx >> 2) & 1) | ((hi >> 1) << 1);
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (row0 + rc * 4 < rows_valid && k0 + k < k_valid)
        v = *reinterpret_cast<const float4*>(g + (size_t)(k0 + k) * col_stride + (row0 + rc * 4));
      const int r = rc * 4;
      float* base = smem + ((r >> 3) * (CH * 32) + (k >> 2) * 32 + (k & 3));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = (jj + (k >> 2)) & 3;
        const float x = j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w));
        base[((r + j) & 7) * 4] = x;
      }
    }
  }
}

template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS) gemm_tf32_kernel(const Params p) {
  extern __shared__ __align__(128) float smem[];
  float* sA = smem;                                // [STAGES][BM * BK]
  float* sB = smem + STAGES * BM * BK;             // [STAGES][BN * BK]
  __shared__ __align__(8) uint64_t bar[STAGES + 1];
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.k_split;
  const int kend = min(p.K, kbeg + p.k_split);

  if (warp == 0) {  // TMEM: BN fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    for (int s = 0; s <= STAGES; ++s) mbar_init(&bar[s], 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_smem;

  // instruction descriptor: D fp32, A / B tf32, both K-major in shared memory (operands that are MN-contiguous in global
  // memory are transposed while being staged: a tf32 MN-major descriptor produced zeros on this part), N >> 3, M >> 4
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  // K-major, no swizzle: core matrices (8 rows x 16 B) 128 B apart along K (LBO), 8-row groups BK / 4 * 128 B apart (SBO);
  // one UMMA_K = 8 step covers two core matrices
  constexpr uint32_t LBO = 128, SBO = (BK / 4) * 128, KSTEP = 2 * 128;

  const int nkb = (kend - kbeg + BK - 1) / BK;
  uint32_t phase[STAGES] = {0, 0};
  for (int kb = 0; kb < nkb; ++kb) {
    const int s = kb & 1;
    if (kb >= STAGES) {  // the MMAs that read this stage two k-blocks ago must have retired
      mbar_wait(&bar[s], phase[s]);
      phase[s] ^= 1;
    }
    const int k0 = kbeg + kb * BK;
    stage_tile<BM, A_MN>(sA + s * BM * BK, p.A, p.a_row, p.a_col, m0, k0, p.M, kend);
    stage_tile<BN, B_MN>(sB + s * BN * BK, p.B, p.b_row, p.b_col, n0, k0, p.N, kend);
    asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes -> visible to the tensor core (async proxy)
    __syncthreads();
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a_addr = smem_u32(sA + s * BM * BK), b_addr = smem_u32(sB + s * BN * BK);
#pragma unroll
      for (int k = 0; k < BK / UMMA_K; ++k) {
        const uint64_t da = make_desc(a_addr + k * KSTEP, LBO, SBO);
        const uint64_t db = make_desc(b_addr + k * KSTEP, LBO, SBO);
        const uint32_t accumulate = (kb > 0 || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
      }
      // arrives on the stage barrier when these (and all earlier) MMAs are done; implies fence::before_thread_sync
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[s])) : "memory");
      if (kb == nkb - 1)
        asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar[STAGES])) : "memory");
    }
  }
  // ---- epilogue: thread = accumulator row (TMEM lane 32 * warp + lane), 16 columns per tcgen05.ld
  mbar_wait(&bar[STAGES], 0);
  asm volatile("tcgen05.fence::after_thread_sync;");
  const int row = m0 + warp * 32 + lane;
  const bool row_ok = row < p.M;
  float* colsum_s = smem;  // [4 warps][BN] (the operand stages are dead now)
  const size_t split_off = p.epilogue == EPI_PARTIAL ? (size_t)blockIdx.z * p.M * p.ldd : 0;
#pragma unroll 1
  for (int c0 = 0; c0 < BN; c0 += 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)(warp * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    float v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = __uint_as_float(r[j]);
    const int col = n0 + c0;
    if (p.epilogue == EPI_BIAS_ACT) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        const float z = v[j] + ((p.bias && col + j < p.N) ? p.bias[col + j] : 0.f);
        if (row_ok && col + j < p.N) {
          if (p.Z) p.Z[(size_t)row * p.ldd + col + j] = z;
          p.D[(size_t)row * p.ldd + col + j] = act_fwd(p.act, z);
        }
      }
    } else if (p.epilogue == EPI_DGRAD) {
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float g = 0.f;
        if (row_ok && col + j < p.N) {
          g = v[j] * act_bwd(p.act, p.act ? p.zprev[(size_t)row * p.ldd + col + j] : 0.f);
          p.D[(size_t)row * p.ldd + col + j] = g;
        }
        if (p.colsum) {  // column sum over this warp's 32 rows (fixed shuffle tree: deterministic)
#pragma unroll
          for (int o = 16; o; o >>= 1) g += __shfl_xor_sync(0xffffffffu, g, o);
          if (lane == 0) colsum_s[warp * BN + c0 + j] = g;
        }
      }
    } else {
#pragma unroll
      for (int j = 0; j < 16; ++j)
        if (row_ok && col + j < p.N) p.D[split_off + (size_t)row * p.ldd + col + j] = v[j];
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (p.epilogue == EPI_DGRAD && p.colsum) {
    for (int c = tid; c < BN; c += THREADS)
      if (n0 + c < p.N)
        p.colsum[(size_t)blockIdx.x * p.ldd + n0 + c] = ((colsum_s[c] + colsum_s[BN + c]) + colsum_s[2 * BN + c]) + colsum_s[3 * BN + c];
  }
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
}

// out[i] = sum_{s < S} in[s * stride + i] (in order: deterministic) for several segments in one launch
struct ReduceSeg { const float* in; float* out; int n, S; long long stride; };
constexpr int MAXSEG = 24;
struct ReduceArgs { ReduceSeg seg[MAXSEG]; int nseg; };
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceArgs a) {
  const ReduceSeg& sg = a.seg[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x) {
    float t = 0.f;
    for (int s = 0; s < sg.S; ++s) t += sg.in[(size_t)s * sg.stride + i];
    sg.out[i] = t;
  }
}

// Backward of the scalar output layer v = h . w + b of the value network (256 -> 1: too thin for the tensor core) fused
// with the activation derivative of the last hidden layer:  dZ[m][j] = g[m] w[j] act'(Z[m][j]);  per-128-row-block
// partials of colsum(dZ) (bias gradient of the last hidden layer), of dw[j] = sum_m g[m] H[m][j] and of db = sum_m g[m].
// thread = column j (coalesced rows), block = 128 rows.
__global__ void __launch_bounds__(256) head_backward_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                           const float* __restrict__ Z, const float* __restrict__ H, int M,
                                                           int n, int ld, int act, float* __restrict__ dZ,
                                                           float* __restrict__ colsum, float* __restrict__ dw_part,
                                                           float* __restrict__ db_part) {
  const int r0 = blockIdx.x * 128, r1 = min(M, r0 + 128);
  for (int j = threadIdx.x; j < n; j += blockDim.x) {
    const float wj = w[j];
    float cs = 0.f, dw = 0.f;
    for (int m = r0; m < r1; ++m) {
      const float gm = g[m];
      const float d = gm * wj * act_bwd(act, act ? Z[(size_t)m * ld + j] : 0.f);
      dZ[(size_t)m * ld + j] = d;
      cs += d;
      dw += gm * H[(size_t)m * ld + j];
    }
    colsum[(size_t)blockIdx.x * ld + j] = cs;
    dw_part[(size_t)blockIdx.x * n + j] = dw;
  }
  if (threadIdx.x == 0) {
    float s = 0.f;
    for (int m = r0; m < r1; ++m) s += g[m];
    db_part[blockIdx.x] = s;
  }
}

template <int BN, bool A_MN, bool B_MN>
inline cudaError_t launch_one(const Params& p, dim3 grid, cudaStream_t stream) {
  const size_t smem = sizeof(float) * STAGES * (BM + BN) * BK;
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    set = true;
  }
  gemm_tf32_kernel<BN, A_MN, B_MN><<<grid, THREADS, smem, stream>>>(p);
  return cudaGetLastError();
}

// mode 0 forward (A, B K-major), 1 dgrad (A K-major, B MN-major), 2 wgrad (A, B MN-major)
inline cudaError_t launch(const Params& p, int mode, cudaStream_t stream) {
  constexpr int BN = 64;  // 128 x 64 tiles: 88 CTAs for a 2816 x 256 output
  const int splits = (p.K + p.k_split - 1) / p.k_split;
  const dim3 grid((p.M + BM - 1) / BM, (p.N + BN - 1) / BN, splits);
  if (mode == 0) return launch_one<BN, false, false>(p, grid, stream);
  if (mode == 1) return launch_one<BN, false, true>(p, grid, stream);
  return launch_one<BN, true, true>(p, grid, stream);
}

}  // namespace gemm
}  // namespace rsrx
