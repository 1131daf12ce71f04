/*
 * rs_model.h -- the flat "compiled model" descriptor.
 *
 * One MJCF robot (plus the stadium floor) compiled ONCE on the host into plain
 * numbers: tree topology, joint frames, link mass/inertia/COM frames, collision
 * geoms, self-collision pair list, solver constants and the gym task constants.
 * Every importer-derived constant of the reference's Bullet dependency is DATA in
 * this struct, not code.  It replaces what the reference
 * obtains at run time through
 *   b3LoadMJCFCommandInit / b3GetJointInfo / b3GetBodyInfo
 *   (roboschool/cpp-household/physics-bullet.cpp:95-131, 206-251)
 * and the per-env class attributes of roboschool/gym_mujoco_walkers.py:14-136.
 *
 * Plain C, fixed capacity, doubles + ints only, so that ctypes, the CUDA library
 * and the CPU oracle all read the same bytes.
 */
#ifndef RS_MODEL_H
#define RS_MODEL_H

#ifdef __cplusplus
extern "C" {
#endif

#define RS_MAX_LINKS 24
#define RS_MAX_GEOMS 24
#define RS_MAX_PAIRS 192
#define RS_MAX_ACT   24
#define RS_MAX_FEET  8
#define RS_MAX_PARTS 32

/* joint types (btMultibodyLink::eFeatherstoneJointType subset) */
#define RS_JOINT_FIXED     0
#define RS_JOINT_REVOLUTE  1
#define RS_JOINT_PRISMATIC 2

/* geom types */
#define RS_GEOM_SPHERE  0
#define RS_GEOM_CAPSULE 1
#define RS_GEOM_BOX     2 /* btBoxShape: g_p0 = centre, g_p1 = half extents (margin included), g_quat = orientation, g_radius = 0;
                             collides with the floor plane only (box pairs are dropped by the model compiler)            */

/* task kinds */
#define RS_TASK_WALKER          0 /* RoboschoolForwardWalker family (gym_forward_walker.py)                  */
#define RS_TASK_PENDULUM        1 /* RoboschoolInvertedPendulum[Swingup] (gym_pendulums.py:70-127)            */
#define RS_TASK_DOUBLE_PENDULUM 2 /* RoboschoolInvertedDoublePendulum (gym_pendulums.py:7-65)                 */
#define RS_TASK_REACHER         3 /* RoboschoolReacher (gym_reacher.py): task_dof = joint0, joint1, target_x, target_y;
                                     task_link = fingertip, task_link2 = target                               */

/* alive rules (gym_mujoco_walkers.py alive_bonus variants) */
#define RS_ALIVE_Z_AND_PITCH 0 /* Hopper/Walker2d: +1 if z>alive_z and |pitch|<1 else -1 */
#define RS_ALIVE_Z           1 /* Ant (+1), Humanoid (+2): bonus if z>alive_z else -1     */
#define RS_ALIVE_ALWAYS      2 /* pendulum-style plumbing envs: never done from alive      */
#define RS_ALIVE_CHEETAH     3 /* HalfCheetah: +1 if |pitch|<1 and none of feet 1,2,4,5 (shins, thighs) touched the
                                  floor in the PREVIOUS step, else -1 (gym_mujoco_walkers.py:50-52)            */

typedef struct rs_model_desc {
    /* ---- sizes ---- */
    int n_links;      /* links excluding the base                                       */
    int n_dof;        /* joint dofs (excluding the 6 base dofs)                          */
    int fixed_base;   /* 1: base is a massless static anchor (Hopper, pendulum)          */
    int n_geoms;
    int n_pairs;      /* self-collision geom pairs                                       */
    int n_act;        /* action dim A                                                    */
    int n_feet;
    int n_parts;      /* named link parts entering body_xyz mean (gym_forward_walker.py:52-54) */
    int obs_dim;      /* 8 + 2A + n_feet                                                 */
    int has_floor;
    int n_iter;       /* PGS iterations (physics-bullet.cpp:370)                         */
    int frame_skip;   /* sub-steps per env step                                          */
    int alive_rule;
    int body_link;    /* link whose pose/speed is robot_body (-1 = base)                 */
    int max_episode_steps;
    int max_contacts; /* capacity of the per-env live contact-point pool (points beyond it are dropped, in manifold order) */

    /* ---- base ---- */
    double base_mass;
    double base_inertia[3];

    /* ---- links (index i, parent[i] in [-1, i-1], -1 = base) ---- */
    int    parent[RS_MAX_LINKS];
    int    jtype[RS_MAX_LINKS];
    int    dof_idx[RS_MAX_LINKS];        /* -1 for fixed joints                           */
    int    has_limit[RS_MAX_LINKS];
    double axis[RS_MAX_LINKS][3];        /* joint axis in link frame (unit)               */
    double zero_rot[RS_MAX_LINKS][4];    /* quat x,y,z,w rotating parent-frame coords into link-frame coords at q=0 */
    double e_vec[RS_MAX_LINKS][3];       /* parent COM -> pivot, parent frame             */
    double d_vec[RS_MAX_LINKS][3];       /* pivot -> link COM, link frame                 */
    double mass[RS_MAX_LINKS];
    double inertia[RS_MAX_LINKS][3];     /* diagonal, COM frame                           */
    double lim_lo[RS_MAX_LINKS];
    double lim_hi[RS_MAX_LINKS];
    double jdamping[RS_MAX_LINKS];
    double max_vel[RS_MAX_LINKS];        /* b3JointInfo.m_jointMaxVelocity (0 for MJCF)   */
    double break_thresh[RS_MAX_LINKS+1]; /* contact breaking threshold, [0]=base, [i+1]=link i */

    /* ---- collision geoms ---- */
    int    g_link[RS_MAX_GEOMS];         /* -1 = base                                     */
    int    g_type[RS_MAX_GEOMS];
    int    g_floor[RS_MAX_GEOMS];        /* collides with the floor plane                 */
    int    g_pad[RS_MAX_GEOMS];
    double g_p0[RS_MAX_GEOMS][3];        /* sphere centre / capsule end 0, link COM frame */
    double g_p1[RS_MAX_GEOMS][3];        /* capsule end 1                                 */
    double g_radius[RS_MAX_GEOMS];
    double g_friction[RS_MAX_GEOMS];
    int    pair_a[RS_MAX_PAIRS];
    int    pair_b[RS_MAX_PAIRS];
    double floor_friction;

    /* ---- world / solver constants ---- */
    double gravity;          /* float32(9.8) squeezed through the boundary (python-binding.cpp:297) */
    double timestep;         /* float32(0.0165/4)                                          */
    double lin_damping;      /* btMultiBody m_linearDamping                                */
    double ang_damping;
    double max_coord_vel;    /* btMultiBody m_maxCoordinateVelocity                        */
    double erp;              /* btContactSolverInfo m_erp                                  */
    double erp2;             /* m_erp2 <- contactERP 0.9 (physics-bullet.cpp:371)          */
    double split_thresh;     /* m_splitImpulsePenetrationThreshold                         */
    double max_limit_impulse;/* btMultiBodyConstraint m_maxAppliedImpulse                  */
    double use_gyro;         /* 1.0: gyroscopic term on                                    */

    /* ---- task constants (gym_forward_walker.py / gym_mujoco_walkers.py) ---- */
    int    act_dof[RS_MAX_ACT];          /* joint dof driven by action k                   */
    double act_gain[RS_MAX_ACT];         /* torque = gain * clip(a,-1,1)                   */
    int    foot_link[RS_MAX_FEET];
    int    part_link[RS_MAX_PARTS];
    double alive_z;
    double alive_bonus;                  /* +1 / +2                                        */
    double electricity_cost;
    double stall_torque_cost;
    double joints_at_limit_cost;
    double foot_collision_cost;
    double initial_z;                    /* <0: latch first observed z (gym_forward_walker.py:58-59) */
    double walk_target_x, walk_target_y;
    double dt_env;                       /* scene.dt = timestep*frame_skip (double, scene_abstract.py:22) */
    /* reset distribution (gym_forward_walker.py:24-30, gym_mujoco_walkers.py:106-128) */
    double reset_joint_spread;           /* 0.1                                            */
    double reset_yaw_spread;             /* pi/16 for Humanoid, 0 otherwise                */
    double reset_base_pos[3];
    double reset_base_quat[4];
    int    reset_set_base;               /* Humanoid: set_pose_and_speed at reset          */
    int    pad1;
    /* moving walk target (RoboschoolHumanoidFlagrun, gym_humanoid_flagrun.py:6-45) */
    int    flag_task;                    /* 0: fixed walk target; 1: flag repositioned when reached / timed out */
    int    flag_timeout_steps;           /* 600/frame_skip env steps (gym_humanoid_flagrun.py:35)               */
    double flag_halflen;                 /* StadiumScene.stadium_halflen   (scene_stadium.py:6)                 */
    double flag_halfwidth;               /* StadiumScene.stadium_halfwidth (scene_stadium.py:7)                 */
    /* task kind: which obs/reward/done epilogue runs after the sub-steps */
    int    task_kind;                    /* RS_TASK_*                                                            */
    int    task_swingup;                 /* InvertedPendulumSwingup: reward cos(theta), never done (gym_pendulums.py:107-112) */
    int    task_link;                    /* InvertedDoublePendulum: link "pole2" whose pose enters obs/reward   */
    int    n_reset;                      /* joints randomised at reset: q = center + U(-spread, spread), draw k */
    int    reset_dof[RS_MAX_ACT];
    int    task_dof[4];                  /* pendulums: slider, hinge, hinge2 dofs                                */
    double reset_center[RS_MAX_ACT];
    double reset_spread[RS_MAX_ACT];     /* per reset joint: q = center + U(-spread, spread)                       */
    int    task_link2;                   /* Reacher: link "target"                                                */
    int    pad2;
    double g_quat[RS_MAX_GEOMS][4];      /* box geoms: orientation in the link COM frame, quaternion x,y,z,w (identity otherwise) */
    /* contact rows (btMultiBodyConstraintSolver::setupMultiBodyContactConstraint): positional term -pen*erp/dt with
       erp = c_erp above c_split_thresh, c_erp2 below it where the term is dropped (split impulse is not solved for
       multibodies); c_split_thresh = -1e30 keeps the term for every penetration.  erp / erp2 / split_thresh above
       govern the joint-limit rows (btMultiBodyJointLimitConstraint::createConstraintRows). */
    double c_erp;
    double c_erp2;
    double c_split_thresh;
    double armature[RS_MAX_LINKS];       /* added to the joint-space inertia D = S^T I^A S of the link's joint (0: not honoured) */
    double stiffness[RS_MAX_LINKS];      /* joint spring torque -k q, applied like jdamping (0: not honoured)                   */
    /* FlagrunHarder's thrown cube (gym_humanoid_flagrun.py:81-118, models_household/cube.urdf): a second free body per env,
       btMultiBody with no links, loaded next to the robot; state = 13 more floats appended to the env state (pos3, quat4
       x,y,z,w cube->world, omega3, vel3, world frame); contact manifolds cube-floor and cube-vs-every robot geom */
    int    has_cube;
    int    pad3;
    double cube_half[3];                 /* btBoxShape half extents (URDF box size / 2)                                          */
    double cube_mass;                    /* cube.urdf:22                                                                         */
    double cube_inertia[3];              /* diagonal; recomputed from the collision shape's AABB (no URDF_USE_INERTIA_FROM_FILE) */
    double cube_friction;                /* lateral friction of the link (UrdfContactInfo default 0.5: cube.urdf has no <contact>) */
    double cube_break_thresh;            /* contact breaking threshold: 0.02 x bounding-sphere disc of the cube's compound shape */
    double cube_start[3];                /* pose at load (gym_humanoid_flagrun.py:84-85)                                         */
} rs_model_desc;

#ifdef __cplusplus
}
#endif
#endif /* RS_MODEL_H */
