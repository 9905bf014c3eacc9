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
// Structure (deliberately plain: these GEMMs are 0.4 GFLOP, 44-88 CTAs, one wave, latency-bound):
//   CTA = 512 threads, tile 128 x BN (BN = 64), accumulators in TMEM (BN fp32 columns).  The whole contraction slice
//   (<= 256) of both operands is staged at once (192 KB of shared memory), so everything a CTA reads is in flight together.
//   Operands that are contiguous along the contraction are loaded by TMA (cp.async.bulk.tensor.2d, one 32-float slab of
//   all rows per copy, 128-byte swizzle, completion on an mbarrier by expect_tx); the trainers make EVERY operand
//   contraction-contiguous by having the producing epilogue (and the Adam step, for the weights) also write the
//   transposed copy, so the forward, dgrad and wgrad GEMMs are all the <.., false, false> instantiation.  An operand
//   that is only available contiguous along M/N is staged by the threads with an in-flight transposition into the
//   canonical no-swizzle UMMA layout (8 x 16-byte core matrices, K-major) — MN-major tf32 shared-memory descriptors
//   produced zeros on this part, so they are not used.
//   fence.proxy.async, one elected thread issues up to 32 x tcgen05.mma.cta_group::1.kind::tf32 (K = 8 each) and
//   tcgen05.commit's them to an mbarrier (bounded wait, __trap on time-out).  Epilogue: tcgen05.ld 32x32b.x16 (thread =
//   accumulator row) -> padded shared memory -> fused bias / activation / activation-derivative / column sums with
//   row-contiguous 16-byte global accesses, and optionally the transposed copy of the result through the same tile.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <algorithm>
#include <cstring>

#include "rsrx_pdl.cuh"

namespace rsrx {
namespace gemm {

// 16 warps per CTA: the tensor-core part needs one thread, but staging the operands and the fused epilogue are SIMT work
// whose latency shrinks with the number of warps an SM has to interleave (4 warps: 15 us per launch, 16 warps: see profiles/)
constexpr int BM = 128, BK = 32, THREADS = 512;
constexpr int UMMA_K = 8;  // tf32

enum Epilogue { EPI_BIAS_ACT = 0, EPI_DGRAD = 1, EPI_PARTIAL = 2 };
enum Act { ACT_NONE = 0, ACT_SILU = 1, ACT_RELU = 2 };

struct Params {
  // TMA descriptors of the operands that are contiguous along the contraction (box 32 floats x 128 | BN rows, 128-byte
  // swizzle); unused (zeroed) for operands that are transposed while being staged
  alignas(64) CUtensorMap tmA;
  alignas(64) CUtensorMap tmB;
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
  float* DT; int ldt;                 // optional transposed copy of D: DT[col][row] ([N][ldt]) — the K-contiguous operand of
                                      // the weight-gradient GEMM, written with row-contiguous (coalesced) stores
  int kcap;                           // contraction length staged at once (multiple of 32, <= KMAX): sizes the shared memory
  unsigned long long* dbg_t;          // profiling aid: 8 globaltimer stamps of CTA 0 (nullptr = off)
};
__device__ __forceinline__ void stamp(const Params& p, int i) {
  if (p.dbg_t && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    p.dbg_t[i] = t;
  }
}

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

// The sigmoid of the fused epilogues uses the hardware exponential and reciprocal (ex2.approx / rcp.approx, ~2 ulp): the
// activation math of a 128 x 64 tile was 6.4 K warp-instructions with expf + IEEE division, as long as the MMAs and the
// stores together, and its inputs carry TF32 rounding (2^-11) anyway.
__device__ __forceinline__ float fast_sigmoid(float z) { return __fdividef(1.f, 1.f + __expf(-z)); }
__device__ __forceinline__ float act_fwd(int act, float z) {
  if (act == ACT_SILU) return z * fast_sigmoid(z);
  if (act == ACT_RELU) return z > 0.f ? z : 0.f;
  return z;
}
__device__ __forceinline__ float act_bwd(int act, float z) {
  if (act == ACT_SILU) { const float s = fast_sigmoid(z); return s * (1.f + z * (1.f - s)); }
  if (act == ACT_RELU) return z > 0.f ? 1.f : 0.f;
  return 1.f;
}

// shared-memory (matrix) descriptor, no swizzle: start >> 4 | LBO >> 4 << 16 | SBO >> 4 << 32 | version 1 << 46
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | ((uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32) | (1ull << 46);
}

// Stage one operand tile (ROWS x KB, row index = the operand's M/N index, column = contraction index) whose GLOBAL layout is
// contiguous along the row index (not the contraction) into shared memory, transposing on the way in.  (Operands that
// are contiguous along the contraction are loaded by TMA instead, see the kernel.)
// The shared-memory image is K-major, no swizzle: core matrix = 8 rows x 16 B; element (r, k) at byte
//       (r % 8) * 16 + (r / 8) * SBO + (k / 4) * 128 + (k % 4) * 4,  SBO = KB / 4 * 128, LBO = 128.
// 16-byte loads of 4 rows at one contraction index, 4 scalar stores, issued in an order rotated by (k >> 2) & 3 so that the
// 32 lanes of one store instruction hit 32 distinct banks.  Lane bits [0:1] k & 3, [2] chunk parity, [3:4] (k >> 2) & 3.
// Rows >= rows_valid and contraction indices >= k_valid are zero-filled.  The contiguous stride is 1 and every row / k
// start is 16-byte aligned (host-checked).  kb: contraction length of this block (multiple of 32, <= KMAX).
template <int ROWS>
__device__ __forceinline__ void stage_tile_transposed(float* smem, const float* __restrict__ g, int col_stride, int row0, int k0,
                                                      int rows_valid, int k_valid, int kb) {
  const int tid = threadIdx.x;
  const int CH = kb >> 2;  // 16-byte chunks per row
  constexpr int RC = ROWS / 4;
  constexpr int U = 8;        // loads in flight per thread
  const int total = RC * kb;  // (4-row chunk, k) items
  for (int base = tid; base < total; base += THREADS * U) {
    float4 v[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * THREADS, hi = idx >> 5;
      const int k = (idx & 3) | (((idx >> 3) & 3) << 2) | ((hi % (kb >> 4)) << 4);
      const int rc = ((idx >> 2) & 1) | ((hi / (kb >> 4)) << 1);
      v[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < total && row0 + rc * 4 < rows_valid && k0 + k < k_valid)
        v[u] = __ldg(reinterpret_cast<const float4*>(g + (size_t)(k0 + k) * col_stride + (row0 + rc * 4)));
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * THREADS, hi = idx >> 5;
      if (idx >= total) break;
      const int k = (idx & 3) | (((idx >> 3) & 3) << 2) | ((hi % (kb >> 4)) << 4);
      const int r = (((idx >> 2) & 1) | ((hi / (kb >> 4)) << 1)) * 4;
      float* base_p = smem + ((r >> 3) * (CH * 32) + (k >> 2) * 32 + (k & 3));
#pragma unroll
      for (int jj = 0; jj < 4; ++jj) {
        const int j = (jj + (k >> 2)) & 3;
        const float x = j == 0 ? v[u].x : (j == 1 ? v[u].y : (j == 2 ? v[u].z : v[u].w));
        base_p[((r + j) & 7) * 4] = x;
      }
    }
  }
}

constexpr int KMAX = 256;  // contraction length held in shared memory at once (A 128 KB + B 64 KB at BN = 64)

// shared-memory descriptor of a K-major tile in the 128-byte-swizzle layout TMA writes: rows 128 B apart, 16-byte chunk
// index XOR (row & 7), 8-row groups 1024 B apart (SBO); LBO is not used inside one swizzle span; layout type 2
__device__ __forceinline__ uint64_t make_desc_sw128(uint32_t saddr) {
  return (uint64_t)((saddr >> 4) & 0x3fff) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// A_MN / B_MN: the operand is contiguous along its row index in global memory -> staged by the threads with an in-flight
// transposition (no-swizzle image).  Otherwise (contiguous along the contraction) -> loaded by TMA (cp.async.bulk.tensor)
// into the 128-byte-swizzled image, one 32-float slab of all rows per copy.
template <int BN, bool A_MN, bool B_MN>
__global__ void __launch_bounds__(THREADS) gemm_tf32_kernel(const __grid_constant__ Params p) {
  extern __shared__ __align__(1024) float smem_raw[];
  // 1024-byte alignment for the swizzled slabs (the launch adds 1 KB of slack)
  float* smem = reinterpret_cast<float*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  float* sA = smem;                 // [BM x kcap]
  float* sB = smem + BM * p.kcap;   // [BN x kcap]
  __shared__ __align__(8) uint64_t bar;
  __shared__ __align__(8) uint64_t tma_bar;
  __shared__ uint32_t tmem_base_smem;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const int kbeg = blockIdx.z * p.k_split;
  const int kend = min(p.K, kbeg + p.k_split);
  pdl::launch_dependents();
  stamp(p, 0);

  if (warp == 0) {  // TMEM: BN fp32 accumulator columns
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(BN));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
  }
  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_init(&tma_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;");
    if (!A_MN) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmA)) : "memory");
    if (!B_MN) asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&p.tmB)) : "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;");
  const uint32_t tmem = tmem_base_smem;
  pdl::wait();  // everything above (TMEM allocation, barriers, descriptor prefetch) overlaps the predecessor's tail
  stamp(p, 1);

  // instruction descriptor: D fp32, A / B tf32, both K-major in shared memory (operands that are MN-contiguous in global
  // memory are transposed while being staged: a tf32 MN-major descriptor produced zeros on this part), N >> 3, M >> 4
  constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

  // The whole contraction slice (<= KMAX) is staged at once — every global load of the CTA is in flight together, one
  // round trip instead of one per k-block — then 4 .. 32 back-to-back MMAs (K = 8 each) and one commit.  Longer
  // contractions repeat the cycle (not on the value-network path: K <= 256, wgrad is split into 256-row slices).
  uint32_t phase = 0, tma_phase = 0;
  bool first = true;
  for (int k0 = kbeg; k0 < kend; k0 += p.kcap) {
    const int kb = min(p.kcap, ((kend - k0) + BK - 1) / BK * BK);
    if (!first) {  // the previous block's MMAs must have retired before its operands are overwritten
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    if ((!A_MN || !B_MN) && tid == 0) {
      // TMA: one copy per 32-float slab and operand, all in flight at once; out-of-range rows / columns arrive as zeros
      const int nslab = kb / BK;
      const uint32_t bytes = (uint32_t)nslab * BK * 4u * ((A_MN ? 0u : (uint32_t)BM) + (B_MN ? 0u : (uint32_t)BN));
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&tma_bar)), "r"(bytes) : "memory");
      for (int sl = 0; sl < nslab; ++sl) {
        if (!A_MN)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(smem_u32(sA + sl * BM * BK)), "l"(reinterpret_cast<uint64_t>(&p.tmA)), "r"(k0 + sl * BK), "r"(m0),
                         "r"(smem_u32(&tma_bar)) : "memory");
        if (!B_MN)
          asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                       ::"r"(smem_u32(sB + sl * BN * BK)), "l"(reinterpret_cast<uint64_t>(&p.tmB)), "r"(k0 + sl * BK), "r"(n0),
                         "r"(smem_u32(&tma_bar)) : "memory");
      }
    }
    if (A_MN) stage_tile_transposed<BM>(sA, p.A, p.a_col, m0, k0, p.M, kend, kb);
    if (B_MN) stage_tile_transposed<BN>(sB, p.B, p.b_col, n0, k0, p.N, kend, kb);
    stamp(p, 2);
    if (A_MN || B_MN) asm volatile("fence.proxy.async.shared::cta;");  // generic-proxy writes -> visible to the tensor core
    __syncthreads();
    if (!A_MN || !B_MN) {
      mbar_wait(&tma_bar, tma_phase);
      tma_phase ^= 1;
    }
    stamp(p, 3);
    if (tid == 0) {
      asm volatile("tcgen05.fence::after_thread_sync;");
      const uint32_t a_addr = smem_u32(sA), b_addr = smem_u32(sB);
      // transposed operands: K-major, no swizzle: core matrices (8 rows x 16 B) 128 B apart along K (LBO), 8-row groups
      // kb / 4 * 128 B apart (SBO), one UMMA_K = 8 step = two core matrices.  TMA operands: slab k / 4, 32 B per step inside
      // the 128-byte swizzle span.
      const uint32_t SBO = (uint32_t)(kb >> 2) * 128;
      // descriptors advance by adding to the 14-bit (address >> 4) field: one 64-bit add per operand per MMA instead of
      // rebuilding them (the issue loop, not the tensor core, was the limit: 45 ns per 128 x 64 x 8 instruction)
      uint64_t da = A_MN ? make_desc(a_addr, 128, SBO) : make_desc_sw128(a_addr);
      uint64_t db = B_MN ? make_desc(b_addr, 128, SBO) : make_desc_sw128(b_addr);
      // byte steps: inside a 32-float slab / from the last step of a slab to the next slab
      constexpr uint32_t A_IN = A_MN ? 256 : 32, A_OUT = A_MN ? 256 : BM * BK * 4 - 3 * 32;
      constexpr uint32_t B_IN = B_MN ? 256 : 32, B_OUT = B_MN ? 256 : BN * BK * 4 - 3 * 32;
      const int nk = kb / UMMA_K;
#pragma unroll 4
      for (int k = 0; k < nk; ++k) {
        const uint32_t accumulate = (!first || k > 0) ? 1u : 0u;
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(tmem), "l"(da), "l"(db), "r"(idesc), "r"(accumulate) : "memory");
        const bool last_of_slab = (k & 3) == 3;
        da += (uint64_t)((last_of_slab ? A_OUT : A_IN) >> 4);
        db += (uint64_t)((last_of_slab ? B_OUT : B_IN) >> 4);
      }
      // arrives on the barrier when these (and all earlier) MMAs are done; implies fence::before_thread_sync
      asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(&bar)) : "memory");
    }
    first = false;
  }
  stamp(p, 4);
  mbar_wait(&bar, phase);
  asm volatile("tcgen05.fence::after_thread_sync;");
  stamp(p, 5);

  // ---- epilogue, pass 1: TMEM -> shared memory (thread = accumulator row = TMEM lane)
  constexpr int LDS = BN + 1;       // padded: a warp writes 32 rows of one column without bank conflicts
  float* sC = smem;                 // [BM][LDS] (the operand image is dead now)
  // warp w reads TMEM lanes 32 * (w % 4) .. + 31 (the only ones it may touch) and column chunks w / 4, w / 4 + 4, ...
#pragma unroll 1
  for (int c0 = (warp >> 2) * 16; c0 < BN; c0 += (THREADS / 128) * 16) {
    uint32_t r[16];
    const uint32_t taddr = tmem + ((uint32_t)((warp & 3) * 32) << 16) + c0;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 16; ++j) sC[((warp & 3) * 32 + lane) * LDS + c0 + j] = __uint_as_float(r[j]);
  }
  asm volatile("tcgen05.fence::before_thread_sync;");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(BN));
  stamp(p, 6);

  // ---- pass 2: fused epilogue with row-contiguous global accesses: thread = (row group tid / 16, 4 columns)
  constexpr int TPR = BN / 4;                 // threads per row
  constexpr int RPP = THREADS / TPR;          // rows per pass
  const int cq = (tid % TPR) * 4, rg = tid / TPR;
  const int col = n0 + cq;
  const size_t split_off = p.epilogue == EPI_PARTIAL ? (size_t)blockIdx.z * p.M * p.ldd : 0;
  float cs[4] = {0.f, 0.f, 0.f, 0.f};
  float bias4[4] = {0.f, 0.f, 0.f, 0.f};
  if (p.epilogue == EPI_BIAS_ACT && p.bias)
#pragma unroll
    for (int j = 0; j < 4; ++j) bias4[j] = col + j < p.N ? p.bias[col + j] : 0.f;
  const bool vec = col + 3 < p.N;  // ldd % 4 == 0 and 16-byte aligned bases (host-checked): whole float4 in range
  // BM / RPP = 4 passes, unrolled: the global loads of the dgrad epilogue (zprev) are all in flight at once
#pragma unroll
  for (int rr = rg; rr < BM; rr += RPP) {
    const int row = m0 + rr;
    if (row >= p.M) break;
    float v[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) v[j] = sC[rr * LDS + cq + j];
    const size_t off = (size_t)row * p.ldd + col;
    if (p.epilogue == EPI_BIAS_ACT) {
      float z[4], y[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) { z[j] = v[j] + bias4[j]; y[j] = act_fwd(p.act, z[j]); }
      if (vec) {
        if (p.Z) *reinterpret_cast<float4*>(p.Z + off) = make_float4(z[0], z[1], z[2], z[3]);
        *reinterpret_cast<float4*>(p.D + off) = make_float4(y[0], y[1], y[2], y[3]);
      } else {
        for (int j = 0; j < 4; ++j)
          if (col + j < p.N) { if (p.Z) p.Z[off + j] = z[j]; p.D[off + j] = y[j]; }
      }
      if (p.DT)
#pragma unroll
        for (int j = 0; j < 4; ++j) sC[rr * LDS + cq + j] = y[j];
    } else if (p.epilogue == EPI_DGRAD) {
      float zp[4] = {0.f, 0.f, 0.f, 0.f}, gg[4];
      if (p.act) {
        if (vec) { const float4 t = *reinterpret_cast<const float4*>(p.zprev + off); zp[0] = t.x; zp[1] = t.y; zp[2] = t.z; zp[3] = t.w; }
        else for (int j = 0; j < 4; ++j) if (col + j < p.N) zp[j] = p.zprev[off + j];
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) { gg[j] = (col + j < p.N) ? v[j] * act_bwd(p.act, zp[j]) : 0.f; cs[j] += gg[j]; }
      if (vec) *reinterpret_cast<float4*>(p.D + off) = make_float4(gg[0], gg[1], gg[2], gg[3]);
      else for (int j = 0; j < 4; ++j) if (col + j < p.N) p.D[off + j] = gg[j];
      if (p.DT)
#pragma unroll
        for (int j = 0; j < 4; ++j) sC[rr * LDS + cq + j] = gg[j];
    } else {
      if (vec) *reinterpret_cast<float4*>(p.D + split_off + off) = make_float4(v[0], v[1], v[2], v[3]);
      else for (int j = 0; j < 4; ++j) if (col + j < p.N) p.D[split_off + off + j] = v[j];
    }
  }
  if (p.DT && p.epilogue != EPI_PARTIAL) {  // pass 3: the transposed copy, a warp per column, lanes over 32 consecutive rows
    __syncthreads();
#pragma unroll 4
    for (int task = warp; task < BN * (BM / 32); task += THREADS / 32) {
      const int c = task / (BM / 32), rr = (task % (BM / 32)) * 32 + lane;
      if (n0 + c < p.N && m0 + rr < p.M) p.DT[(size_t)(n0 + c) * p.ldt + m0 + rr] = sC[rr * LDS + c];
    }
  }
  stamp(p, 7);
  if (p.epilogue == EPI_DGRAD && p.colsum) {  // column sums over the CTA's rows: row groups combined in a fixed order
    __syncthreads();
    float* red = smem + BM * LDS;  // [RPP][BN]
#pragma unroll
    for (int j = 0; j < 4; ++j) red[rg * BN + cq + j] = cs[j];
    __syncthreads();
    for (int c = tid; c < BN; c += THREADS) {
      float t = 0.f;
      for (int g2 = 0; g2 < RPP; ++g2) t += red[g2 * BN + c];
      if (n0 + c < p.N) p.colsum[(size_t)blockIdx.x * p.ldd + n0 + c] = t;
    }
  }
}

// out[i] = sum_{s < S} in[s * stride + i] (in order: deterministic) for several segments in one launch
struct ReduceSeg { const float* in; float* out; int n, S; long long stride; };
constexpr int MAXSEG = 24;
struct ReduceArgs { ReduceSeg seg[MAXSEG]; int nseg; };
__global__ void __launch_bounds__(256) reduce_partials_kernel(const ReduceArgs a) {
  pdl::launch_dependents();
  pdl::wait();
  const ReduceSeg& sg = a.seg[blockIdx.y];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x) {
    float t = 0.f;
    int s = 0;
    for (; s + 8 <= sg.S; s += 8) {  // 8 loads in flight; added in slice order
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = sg.in[(size_t)(s + u) * sg.stride + i];
#pragma unroll
      for (int u = 0; u < 8; ++u) t += v[u];
    }
    for (; s < sg.S; ++s) t += sg.in[(size_t)s * sg.stride + i];
    sg.out[i] = t;
  }
}

// Backward of the scalar output layer v = h . w + b of the value network (256 -> 1: too thin for the tensor core) fused
// with the activation derivative of the last hidden layer:  dZ[m][j] = g[m] w[j] act'(Z[m][j]);  per-128-row-block
// partials of colsum(dZ) (bias gradient of the last hidden layer), of dw[j] = sum_m g[m] H[m][j] and of db = sum_m g[m].
// Block = 128 rows x 64 columns: 16 x 16 threads, a thread owns 4 columns (16-byte accesses) of 8 rows; the 16 row lanes
// are combined through shared memory in a fixed order (deterministic).  n % 4 == 0, ld % 4 == 0.
__global__ void __launch_bounds__(256) head_backward_kernel(const float* __restrict__ g, const float* __restrict__ w,
                                                           const float* __restrict__ Z, const float* __restrict__ H, int M,
                                                           int n, int ld, int act, float* __restrict__ dZ,
                                                           float* __restrict__ colsum, float* __restrict__ dw_part,
                                                           float* __restrict__ db_part, float* __restrict__ dZT, int ldt) {
  __shared__ float red[2][16][64];
  __shared__ float tile[128][65];  // for the transposed copy dZT[col][row]
  pdl::launch_dependents();
  pdl::wait();
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int r0 = blockIdx.x * 128, col = blockIdx.y * 64 + tx * 4;
  float cs[4] = {0.f, 0.f, 0.f, 0.f}, dw[4] = {0.f, 0.f, 0.f, 0.f};
  if (col < n) {
    const float4 w4 = *reinterpret_cast<const float4*>(w + col);
    const float wj[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int m = r0 + ty + 16 * i;
      if (m < M) {
        const float gm = g[m];
        const size_t off = (size_t)m * ld + col;
        const float4 h4 = *reinterpret_cast<const float4*>(H + off);
        float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (act) z4 = *reinterpret_cast<const float4*>(Z + off);
        const float zz[4] = {z4.x, z4.y, z4.z, z4.w}, hh[4] = {h4.x, h4.y, h4.z, h4.w};
        float d[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          d[j] = gm * wj[j] * act_bwd(act, zz[j]);
          cs[j] += d[j];
          dw[j] += gm * hh[j];
          tile[ty + 16 * i][tx * 4 + j] = d[j];
        }
        *reinterpret_cast<float4*>(dZ + off) = make_float4(d[0], d[1], d[2], d[3]);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) { red[0][ty][tx * 4 + j] = cs[j]; red[1][ty][tx * 4 + j] = dw[j]; }
  __syncthreads();
  if (dZT) {  // a warp per column, lanes over 32 consecutive rows: coalesced
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int task = warp; task < 64 * 4; task += 8) {
      const int c = task >> 2, rr = (task & 3) * 32 + lane, cc = blockIdx.y * 64 + c;
      if (cc < n && r0 + rr < M) dZT[(size_t)cc * ldt + r0 + rr] = tile[rr][c];
    }
  }
  if (threadIdx.x < 128) {
    const int which = threadIdx.x >> 6, c = threadIdx.x & 63;
    float t = 0.f;
    for (int y = 0; y < 16; ++y) t += red[which][y][c];
    const int cc = blockIdx.y * 64 + c;
    if (cc < n) {
      if (which == 0) colsum[(size_t)blockIdx.x * ld + cc] = t;
      else dw_part[(size_t)blockIdx.x * n + cc] = t;
    }
  }
  if (blockIdx.y == 0 && threadIdx.x >= 128 && threadIdx.x < 160) {  // one warp: db partial of this row block
    const int lane = threadIdx.x - 128;
    float sgm = 0.f;
    for (int m = r0 + lane; m < min(M, r0 + 128); m += 32) sgm += g[m];
#pragma unroll
    for (int o = 16; o; o >>= 1) sgm += __shfl_xor_sync(0xffffffffu, sgm, o);
    if (lane == 0) db_part[blockIdx.x] = sgm;
  }
}

// Adam on a list of tensors in one launch (torch.optim.Adam semantics, no weight decay / amsgrad): the trainers' networks
// have ~24 small tensors; torch's multi-tensor kernel takes 40 us on them.  The step count lives on the device so that a
// captured CUDA graph advances it on every replay: every block READS the count k at its start (plain load) and bumps a
// `done` counter at its end; the last block to finish publishes k + 1 and clears `done`.  Launches on a stream do not
// overlap, and the count only changes after every block has read it: no race, no second kernel, and no atomic on any
// block's critical path.  counters = {step count, done} (two 32-bit words of the caller's uint64 ticket).
struct AdamSeg { float* p; const float* g; float* m; float* v; int n; float* pT; int cols; };  // pT: optional [cols][n / cols] copy
constexpr int ADAM_MAXSEG = 32;
constexpr int ADAM_BLOCKS_X = 64;  // one 32 x 32 tile per block on the 256 x 256 weights: no serial tile loop
struct AdamArgs { AdamSeg seg[ADAM_MAXSEG]; int nseg; float lr, beta1, beta2, eps, grad_scale; unsigned int* counters; };
__global__ void __launch_bounds__(256) adam_kernel(const AdamArgs a) {
  __shared__ float tile[32][33];
  pdl::launch_dependents();
  pdl::wait();
  const AdamSeg& sg = a.seg[blockIdx.y];
  const unsigned int k_now = *reinterpret_cast<volatile unsigned int*>(a.counters);
  const float t = (float)(k_now + 1u);
  const float bc1 = 1.f - expf(t * logf(a.beta1)), bc2 = 1.f - expf(t * logf(a.beta2));
  const float step_size = a.lr / bc1, inv_sqrt_bc2 = rsqrtf(bc2);
  auto update = [&](int i) {
    const float gr = sg.g[i] * a.grad_scale;
    const float m = a.beta1 * sg.m[i] + (1.f - a.beta1) * gr;
    const float v = a.beta2 * sg.v[i] + (1.f - a.beta2) * gr * gr;
    sg.m[i] = m; sg.v[i] = v;
    const float pn = sg.p[i] - step_size * m / (sqrtf(v) * inv_sqrt_bc2 + a.eps);
    sg.p[i] = pn;
    return pn;
  };
  if (!sg.pT) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < sg.n; i += gridDim.x * blockDim.x) update(i);
  } else {
  // 2-D tensor with a transposed copy (the dgrad GEMM's TMA operand): 32 x 32 tiles through shared memory, so that both
  // the parameter and its transpose are written with contiguous rows
  const int cols = sg.cols, rows = sg.n / cols;
  const int tc = (cols + 31) / 32, tr = (rows + 31) / 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;  // 32 x 8
  for (int tl = blockIdx.x; tl < tc * tr; tl += gridDim.x) {
    const int r0 = (tl / tc) * 32, c0 = (tl % tc) * 32;
    __syncthreads();
    for (int y = ty; y < 32; y += 8)
      if (r0 + y < rows && c0 + tx < cols) tile[y][tx] = update((r0 + y) * cols + c0 + tx);
    __syncthreads();
    for (int y = ty; y < 32; y += 8)
      if (c0 + y < cols && r0 + tx < rows) sg.pT[(size_t)(c0 + y) * rows + r0 + tx] = tile[tx][y];
  }
  }
  // the last block to finish advances the step count
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned int blocks = gridDim.x * gridDim.y;
    if (atomicAdd(a.counters + 1, 1u) == blocks - 1) {
      a.counters[1] = 0u;
      a.counters[0] = k_now + 1u;
      __threadfence();
    }
  }
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link against libcuda)
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  }
  return fn;
}
// row-major fp32 matrix [rows][cols], leading dimension ld: box = 32 columns (128 B, the swizzle span) x box_rows rows
inline cudaError_t make_tmap(CUtensorMap* tm, const float* base, int rows, int cols, int ld, int box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return cudaErrorNotSupported;
  const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)BK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

template <int BN, bool A_MN, bool B_MN>
inline cudaError_t launch_one(Params p, dim3 grid, cudaStream_t stream) {
  const int kslice = std::min(p.K, p.k_split);
  p.kcap = std::min(KMAX, (kslice + BK - 1) / BK * BK);
  const size_t epi = sizeof(float) * (BM * (BN + 1) + (THREADS / (BN / 4)) * BN);  // epilogue staging + column-sum scratch
  const size_t smem = std::max(sizeof(float) * (size_t)(BM + BN) * p.kcap, epi) + 1024;  // + alignment slack
  memset(&p.tmA, 0, sizeof(p.tmA));
  memset(&p.tmB, 0, sizeof(p.tmB));
  if (!A_MN) { cudaError_t e = make_tmap(&p.tmA, p.A, p.M, p.K, p.a_row, BM); if (e != cudaSuccess) return e; }
  if (!B_MN) { cudaError_t e = make_tmap(&p.tmB, p.B, p.N, p.K, p.b_row, BN); if (e != cudaSuccess) return e; }
  static bool set = false;
  if (!set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_tf32_kernel<BN, A_MN, B_MN>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)(sizeof(float) * (BM + BN) * KMAX + 1024));
    if (e != cudaSuccess) return e;
    set = true;
  }
  return pdl::launch(gemm_tf32_kernel<BN, A_MN, B_MN>, grid, dim3(THREADS), smem, stream, p);
}

// mode 0: A and B contiguous along the contraction (both by TMA: forward, dgrad with a transposed weight copy, wgrad on
// transposed activation copies); 1: B contiguous along its rows (dgrad straight from W); 2: both (wgrad from dZ and X)
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
