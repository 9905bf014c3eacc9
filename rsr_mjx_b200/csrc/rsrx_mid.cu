// rsrx_mid.cu — the mid-capacity instantiation of step_kernel (see rsrx_mid.h).  Same source as the fast one
// (rsrx_env.cuh), same arithmetic in the same order: for an env that fits both, the two give the same bits
// (tests/test_gpu_parity.py::test_mid_capacity_kernel_is_bitwise_the_fast_kernel).  The Jacobian-row pool holds every row
// of all 64 contacts, so this instantiation never spills; an env with more than 64 active contacts still goes to the
// redo list and the 256-contact kernel.
#define RSRX_NS rsrx_mid
#define RSRX_MAXC 64
#define RSRX_WPB 8
#define RSRX_STEP_ONLY
#include "rsrx_env.cuh"
#include "rsrx_mid.h"

using namespace rsrx_mid;
static_assert(WPB == kMidWarps, "rsrx_mid.h and rsrx_mid.cu disagree on the CTA shape");

int rsrx_mid_prepare(const void* host_dmodel, size_t bytes, int max_smem_optin, void** dev_out, int* arena_bytes_out,
                     const char** err) {
  if (bytes != sizeof(DModel)) { *err = "rsrx_mid_prepare: DModel size mismatch between the instantiations"; return 1; }
  DModel d;
  memcpy(&d, host_dmodel, sizeof(DModel));
  int stride = (max_smem_optin / (int)sizeof(float) / WPB) & ~3;
  int pool = stride - ar::FIXED;
  if (pool > MAXC * 4 * NCOL) pool = MAXC * 4 * NCOL;
  if (pool < MAXC * 4 * NCOL) { *err = "rsrx_mid_prepare: not enough shared memory per block for the mid arena"; return 1; }
  d.arena_stride = (ar::FIXED + pool + 3) & ~3;  // 16-byte multiples: the arena starts with 8-byte pointers
  d.pool_floats = pool;
  d.contact_cap = MAXC;
  DModel* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, sizeof(DModel));
  if (e == cudaSuccess) e = cudaMemcpy(dev, &d, sizeof(DModel), cudaMemcpyHostToDevice);
  if (e == cudaSuccess)
    e = cudaFuncSetAttribute(step_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, WPB * d.arena_stride * (int)sizeof(float));
  if (e != cudaSuccess) { if (dev) cudaFree(dev); *err = cudaGetErrorString(e); return 1; }
  *dev_out = dev;
  *arena_bytes_out = d.arena_stride * (int)sizeof(float);
  return 0;
}

cudaError_t rsrx_mid_launch_step(const void* dev_dmodel, const rsrx_mid_launch& a, int num_sms, int arena_bytes, cudaStream_t s) {
  // one CTA per SM, as evenly filled as possible (N <= num_sms * WPB is the caller's condition for choosing this kernel)
  int w = (a.N + num_sms - 1) / num_sms;
  w = w < 1 ? 1 : (w > WPB ? WPB : w);
  PerEnv pe = {a.geom_friction, a.body_mass, a.dof_damping, a.dof_frictionloss, nullptr, a.redo};
  StatePtrs st;
  st.data = a.data; st.first_data = a.first_data; st.obs = a.obs; st.first_obs = a.first_obs; st.reward = a.reward;
  st.done = a.done; st.info = a.info; st.metrics = a.metrics; st.status = a.status;
  step_kernel<WPB><<<(a.N + w - 1) / w, 32 * w, (size_t)w * arena_bytes, s>>>(static_cast<const DModel*>(dev_dmodel), a.N, a.action, pe, st);
  return cudaGetLastError();
}
