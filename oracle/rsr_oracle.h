/* rsr_oracle.h — public structs of the CPU oracle (TEST INFRASTRUCTURE ONLY).
 *
 * The interface is float64 whatever precision the oracle computes in
 * (-DORACLE_REAL=float|double); float32 values survive the round trip exactly.
 * Mirrored by oracle/oracle.py (ctypes).
 */
#ifndef RSR_ORACLE_H_
#define RSR_ORACLE_H_
#include "../include/rsrx_model.h"

#define ORC_MAXCON (4 * RSRX_MAXPAIR)
#define ORC_MAXEFC (RSRX_MAXEQ + 2 * RSRX_MAXV + 6 * ORC_MAXCON)
#define ORC_MAXOBS 24

/* mjx.Data subset that callers of the Airbot envs read */
typedef struct orc_data {
  double qpos[RSRX_MAXQ], qvel[RSRX_MAXV], ctrl[RSRX_MAXU],
      qacc_warmstart[RSRX_MAXV], time;
  double xpos[RSRX_MAXBODY][3], xquat[RSRX_MAXBODY][4];
  double site_xpos[RSRX_MAXSITE][3], geom_xpos[RSRX_MAXGEOM][3];
  /* diagnostics of the last forward() */
  double qacc[RSRX_MAXV], qacc_smooth[RSRX_MAXV], qfrc_constraint[RSRX_MAXV],
      qfrc_bias[RSRX_MAXV], qfrc_actuator[RSRX_MAXV];
  int32_t ncon, ncon_active, nefc, nefc_active, solver_niter, ls_total;
  int32_t pad[2];
} orc_data;

/* brax State of the wrapped env (AutoReset(Episode(Vmap(env)))) for one env */
typedef struct orc_env_state {
  orc_data d;     /* pipeline_state */
  orc_data first; /* info['first_pipeline_state'] */
  double obs[ORC_MAXOBS], first_obs[ORC_MAXOBS];
  double reward, done, truncation, steps;
  /* info (union over the three env kinds)
   *  sf/cube: target_pos[3] new_cube_pos[2] site_pos[3] cube_pos[3] last_action reached_box
   *  T:       target_base_pos[3] target_vertical_pos[3] target_w new_T_pos[2]
   *           site_pos[3] T_pos[3] xita */
  double target_pos[3], target2_pos[3], new_pos[2], site_pos[3], obj_pos[3],
      last_action, xita, target_w;
  /* metrics: sf/cube: push_reward ctrl_cost siet_to_box_reward
   *          T: push_reward siet2cube_reward health_reward task_complete_reward site_z_reward */
  double metrics[5];
} orc_env_state;

typedef struct orc_contact {
  double dist, pos[3], frame[9], friction[5], solref[2], solimp[5];
  int32_t geom1, geom2;
} orc_contact;

#endif
