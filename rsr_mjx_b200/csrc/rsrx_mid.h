// rsrx_mid.h — host interface between rsrx_api.cu (namespace rsrx) and rsrx_mid.cu (namespace rsrx_mid: the stepper's
// step_kernel compiled for 64 active contacts per env and at most 8 envs per CTA).  Small batches — up to 8 envs per SM, e.g.
// the 1024 envs of the PPO / SAC trainers — leave most of an SM's shared memory unused, so their arena can afford the
// contact capacity that makes the large-capacity redo pass (one more env-step latency whenever an env exceeds the fast
// arena's 24 contacts: 5-10 % of a PPO collect phase) practically never necessary.  Plain structs only.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct rsrx_mid_launch {
  int N;
  const float* action;
  const float *geom_friction, *body_mass, *dof_damping, *dof_frictionloss;
  int* redo;
  float *data, *first_data, *obs, *first_obs, *reward, *done, *info, *metrics;
  int* status;
};
constexpr int kMidWarps = 8;  // most envs per CTA of the mid-capacity step kernel

// uploads a copy of the host DModel patched with the mid arena; returns 0 on success
int rsrx_mid_prepare(const void* host_dmodel, size_t bytes, int max_smem_optin, void** dev_dmodel_out, int* arena_bytes_out,
                     const char** err);
cudaError_t rsrx_mid_launch_step(const void* dev_dmodel, const rsrx_mid_launch& a, int num_sms, int arena_bytes, cudaStream_t s);
