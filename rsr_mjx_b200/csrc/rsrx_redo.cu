// rsrx_redo.cu — the large-capacity instantiation of the stepper.  Same source as the fast one (rsrx_env.cuh), compiled
// with room for every contact slot of every geom pair (4 x RSRX_MAXPAIR = 256 active contacts per env, what MJX keeps
// densely), two envs per CTA, no phase barriers.  rsrx_env_reset / step / physics_step launch it after the fast kernel
// on the (normally empty) list of envs whose substep exceeded the fast arena's 24 contacts, so that no env ever loses
// a contact.  Same arithmetic in the same order: for an env that does fit, both instantiations give the same bits
// (tests/test_gpu_parity.py::test_redo_path_is_bitwise_identical).
#define RSRX_NS rsrx_big
#define RSRX_MAXC (4 * 64)
#define RSRX_WPB 2
#define RSRX_SYNC_MASK 0
#define RSRX_REDO_ONLY
#include "rsrx_env.cuh"
#include "rsrx_redo.h"

using namespace rsrx_big;
static_assert(RSRX_MAXC == 4 * RSRX_MAXPAIR, "the redo arena must hold every slot of every pair");
static_assert(WPB == kRedoWarps, "rsrx_redo.h and rsrx_redo.cu disagree on the CTA shape");

int rsrx_big_prepare(const void* host_dmodel, size_t bytes, int max_smem_optin, void** dev_out, int* smem_bytes_out,
                     int* max_contacts_out, const char** err) {
  if (bytes != sizeof(DModel)) { *err = "rsrx_big_prepare: DModel size mismatch between the two instantiations"; return 1; }
  DModel d;
  memcpy(&d, host_dmodel, sizeof(DModel));
  if (4 * d.npair > MAXC) { *err = "rsrx_big_prepare: model has more contact slots than the redo arena"; return 1; }
  int stride = (max_smem_optin / (int)sizeof(float) / WPB) & ~3;
  int pool = stride - ar::FIXED;
  if (pool > MAXC * 4 * NCOL) pool = MAXC * 4 * NCOL;
  // the pool holds the Jacobian rows of all MAXC contacts, so this instantiation never spills (and gets no spill buffer)
  if (pool < MAXC * 4 * NCOL) { *err = "rsrx_big_prepare: not enough shared memory per block for the redo arena"; return 1; }
  d.arena_stride = (ar::FIXED + pool + 3) & ~3;  // 16-byte multiples: the arena starts with 8-byte pointers
  d.pool_floats = pool;
  d.contact_cap = MAXC;
  const int smem = WPB * d.arena_stride * (int)sizeof(float);
  DModel* dev = nullptr;
  cudaError_t e = cudaMalloc(&dev, sizeof(DModel));
  if (e == cudaSuccess) e = cudaMemcpy(dev, &d, sizeof(DModel), cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaFuncSetAttribute(redo_kernel<WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) { if (dev) cudaFree(dev); *err = cudaGetErrorString(e); return 1; }
  *dev_out = dev;
  *smem_bytes_out = smem;
  *max_contacts_out = MAXC;
  return 0;
}

cudaError_t rsrx_big_launch(const void* dev_dmodel, const rsrx_redo_launch& a, int grid, int smem_bytes, cudaStream_t s) {
  RedoArgs ra;
  ra.mode = a.mode; ra.qpos = a.qpos; ra.qvel = a.qvel; ra.ctrl = a.ctrl; ra.action = a.action;
  ra.data = a.phys_data; ra.nsteps = a.nsteps; ra.status_out = a.phys_status;
  PerEnv pe = {a.geom_friction, a.body_mass, a.dof_damping, a.dof_frictionloss, nullptr, a.redo};
  StatePtrs st;
  st.data = a.data; st.first_data = a.first_data; st.obs = a.obs; st.first_obs = a.first_obs; st.reward = a.reward;
  st.done = a.done; st.info = a.info; st.metrics = a.metrics; st.status = a.status;
  redo_kernel<WPB><<<grid, 32 * WPB, smem_bytes, s>>>(static_cast<const DModel*>(dev_dmodel), ra, pe, st);
  return cudaGetLastError();
}
