/* rsrx_model.h — compiled-model blob shared by the host compiler, the CUDA
 * library and the CPU oracle.
 *
 * The reference has no native boundary: its model is an `mjx.Model` pytree built
 * by `mujoco.MjModel.from_xml_path` + `brax.io.mjcf.load_model`
 * (reference test/airbot.py:43-47, ppo_train/airbot_training/cube_env.py:37-43,
 * ppo_train/airbot_training/T_shape_env.py:39-45).  This struct is the flat,
 * fixed-capacity equivalent of the subset of `mjModel` fields that the Airbot
 * hot path reads.  Field names follow mjModel.  All reals are float64; the
 * CUDA library narrows to float32 at rsrx_model_create time.
 *
 * The layout is mirrored field-for-field by rsr_mjx_b200/model.py (ctypes);
 * rsrx_model_blob_size() lets the host check the two agree.
 */
#ifndef RSRX_MODEL_H_
#define RSRX_MODEL_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RSRX_MAGIC 0x52535258 /* 'RSRX' */
#define RSRX_VERSION 3

#define RSRX_MAXBODY 16
#define RSRX_MAXJNT 12
#define RSRX_MAXQ 24
#define RSRX_MAXV 20
#define RSRX_MAXU 8
#define RSRX_MAXGEOM 32
#define RSRX_MAXSITE 4
#define RSRX_MAXPAIR 64
#define RSRX_MAXEQ 2

/* mjtJoint */
#define RSRX_JNT_FREE 0
#define RSRX_JNT_SLIDE 2
#define RSRX_JNT_HINGE 3
/* mjtGeom */
#define RSRX_GEOM_PLANE 0
#define RSRX_GEOM_BOX 6

/* environment kinds (which reference env class the fused pre/post-physics
 * logic follows) */
#define RSRX_ENV_SF 0   /* test/airbot.py             (23-d obs) */
#define RSRX_ENV_CUBE 1 /* airbot_training/cube_env.py (23-d obs) */
#define RSRX_ENV_T 2    /* airbot_training/T_shape_env.py (16-d obs) */

typedef struct rsrx_model_blob {
  int32_t magic, version;
  /* sizes */
  int32_t nbody, njnt, nq, nv, nu, ngeom, nsite, npair, neq;
  /* mjOption */
  int32_t iterations, ls_iterations;
  int32_t pad0;
  double timestep, gravity[3], tolerance, ls_tolerance, impratio;
  /* mjStatistic */
  double meaninertia;

  /* bodies */
  int32_t body_parentid[RSRX_MAXBODY], body_rootid[RSRX_MAXBODY],
      body_weldid[RSRX_MAXBODY], body_jntadr[RSRX_MAXBODY],
      body_jntnum[RSRX_MAXBODY], body_dofadr[RSRX_MAXBODY],
      body_dofnum[RSRX_MAXBODY], body_depth[RSRX_MAXBODY];
  double body_pos[RSRX_MAXBODY][3], body_quat[RSRX_MAXBODY][4],
      body_ipos[RSRX_MAXBODY][3], body_iquat[RSRX_MAXBODY][4],
      body_mass[RSRX_MAXBODY], body_inertia[RSRX_MAXBODY][3],
      body_invweight0[RSRX_MAXBODY][2];

  /* joints */
  int32_t jnt_type[RSRX_MAXJNT], jnt_qposadr[RSRX_MAXJNT],
      jnt_dofadr[RSRX_MAXJNT], jnt_bodyid[RSRX_MAXJNT],
      jnt_limited[RSRX_MAXJNT], jnt_actfrclimited[RSRX_MAXJNT];
  double jnt_pos[RSRX_MAXJNT][3], jnt_axis[RSRX_MAXJNT][3],
      jnt_range[RSRX_MAXJNT][2], jnt_actfrcrange[RSRX_MAXJNT][2],
      jnt_solref[RSRX_MAXJNT][2], jnt_solimp[RSRX_MAXJNT][5],
      jnt_margin[RSRX_MAXJNT];
  double qpos0[RSRX_MAXQ];

  /* dofs */
  int32_t dof_bodyid[RSRX_MAXV], dof_jntid[RSRX_MAXV], dof_parentid[RSRX_MAXV];
  double dof_damping[RSRX_MAXV], dof_frictionloss[RSRX_MAXV],
      dof_armature[RSRX_MAXV], dof_invweight0[RSRX_MAXV],
      dof_solref[RSRX_MAXV][2], dof_solimp[RSRX_MAXV][5];

  /* geoms */
  int32_t geom_type[RSRX_MAXGEOM], geom_bodyid[RSRX_MAXGEOM],
      geom_contype[RSRX_MAXGEOM], geom_conaffinity[RSRX_MAXGEOM],
      geom_condim[RSRX_MAXGEOM], geom_priority[RSRX_MAXGEOM];
  double geom_pos[RSRX_MAXGEOM][3], geom_quat[RSRX_MAXGEOM][4],
      geom_size[RSRX_MAXGEOM][3], geom_friction[RSRX_MAXGEOM][3],
      geom_solref[RSRX_MAXGEOM][2], geom_solimp[RSRX_MAXGEOM][5],
      geom_solmix[RSRX_MAXGEOM], geom_margin[RSRX_MAXGEOM],
      geom_gap[RSRX_MAXGEOM];

  /* sites */
  int32_t site_bodyid[RSRX_MAXSITE];
  double site_pos[RSRX_MAXSITE][3], site_quat[RSRX_MAXSITE][4];

  /* collision pairs after MJX's static filter (collision_driver.geom_pairs):
   * geom1/geom2 ordered by geom type, then by body. */
  int32_t pair_geom1[RSRX_MAXPAIR], pair_geom2[RSRX_MAXPAIR];

  /* actuators (joint transmission, fixed gain, affine bias) */
  int32_t act_trnid[RSRX_MAXU], act_ctrllimited[RSRX_MAXU],
      act_forcelimited[RSRX_MAXU];
  double act_gear[RSRX_MAXU], act_gainprm[RSRX_MAXU][3],
      act_biasprm[RSRX_MAXU][3], act_ctrlrange[RSRX_MAXU][2],
      act_forcerange[RSRX_MAXU][2];

  /* equality constraints (type JOINT only) */
  int32_t eq_obj1id[RSRX_MAXEQ], eq_obj2id[RSRX_MAXEQ];
  double eq_data[RSRX_MAXEQ][5], eq_solref[RSRX_MAXEQ][2],
      eq_solimp[RSRX_MAXEQ][5];
} rsrx_model_blob;

/* Environment constants: ids cached by AirbotPlayBase.__init__ and the reward /
 * reset constants of the three reference env classes (test/airbot.py:10-100,
 * cube_env.py:9-94, T_shape_env.py:11-97). */
typedef struct rsrx_env_cfg {
  int32_t env_kind;       /* RSRX_ENV_* */
  int32_t episode_length; /* brax EpisodeWrapper */
  int32_t action_repeat;  /* brax EpisodeWrapper (1 everywhere in the reference) */
  int32_t n_frames;       /* physics substeps per env step (decimation, 4) */
  int32_t cube_body;      /* cube_for_push | T_block */
  int32_t target_body;    /* target_pos | T_target */
  int32_t site_endpoint;  /* 'endpoint' */
  int32_t site_tail;      /* T: 'T_tail' */
  int32_t site_target_tail; /* T: 'T_target_tail' */
  int32_t geom_base, geom_vertical;               /* T: base_block, vertical_block */
  int32_t geom_target_base, geom_target_vertical; /* T: base_target, vertical_target */
  int32_t joint_qadr[6];  /* qpos addresses of joint1..joint6 */
  int32_t pad0;
  double action_scale[RSRX_MAXU];
  double push_reward_weight, siet_to_box_reward_weight, healthy_reward,
      endpoint_min_z_pos;
} rsrx_env_cfg;

#ifdef __cplusplus
}
#endif
#endif /* RSRX_MODEL_H_ */
