// rsrx_redo.h — host interface between rsrx_api.cu (namespace rsrx, the fast instantiation of the stepper) and
// rsrx_redo.cu (namespace rsrx_big, the large-capacity instantiation that re-runs env-steps with more active contacts
// than the fast arena holds).  Plain structs only: the two translation units see differently-sized arenas.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

struct rsrx_redo_launch {
  int mode;  // 0 reset, 1 step, 2 physics
  const float *qpos, *qvel, *ctrl, *action;
  float* phys_data;
  int nsteps;
  int* phys_status;
  const float *geom_friction, *body_mass, *dof_damping, *dof_frictionloss;
  int* redo;
  float *data, *first_data, *obs, *first_obs, *reward, *done, *info, *metrics;
  int* status;
};

// uploads a copy of the host DModel patched with the large arena; returns 0 on success
int rsrx_big_prepare(const void* host_dmodel, size_t bytes, int max_smem_optin, void** dev_dmodel_out, int* smem_bytes_out,
                     int* max_contacts_out, const char** err);
cudaError_t rsrx_big_launch(const void* dev_dmodel, const rsrx_redo_launch& a, int grid, int smem_bytes, cudaStream_t s);
constexpr int kRedoWarps = 2;  // envs per CTA of the redo kernel
