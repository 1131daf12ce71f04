/*
 * roboschool_b200.h -- C ABI of libroboschool_b200.so: the B200-native batched replacement for the
 * reference's cpp_household World::step hot path.
 *
 * Drop-in boundary: the reference binds this path through the Boost.Python module `cpp_household`
 * (roboschool/cpp-household/python-binding.cpp:605-725) over Household::World
 * (roboschool/cpp-household/household.h:138-205).  Each entry point below names the reference
 * interface it replaces.  A pure-Python `cpp_household` shim (roboschool_b200/cpp_household.py)
 * and the batched VecEnv (roboschool_b200/vec_env.py) call these through ctypes; INTEGRATION.md
 * shows the binding a roboschool maintainer would add.
 *
 * Conventions: every function returns 0 on success, non-zero on error with the message available
 * from rs_last_error(); no exceptions cross the ABI; the library owns worlds/policies, the caller
 * owns every buffer it passes; `*_dev` pointers are CUDA device pointers (e.g. torch tensors'
 * data_ptr()), `*_host` pointers are host memory; `stream` is a cudaStream_t (NULL = default
 * stream).  There is NO CPU fallback: without a CUDA device every compute call fails loudly.
 *
 * Stream ordering: entry points that take a `stream` enqueue on it and return without waiting.  The `*_host` entry points,
 * rs_world_substeps, rs_world_link_poses and the host getters / setters run on a private stream of the world and wait for
 * everything the world has enqueued on caller streams before (an event is recorded behind every enqueue), so
 * `rs_world_reset(stream); rs_world_step_host(...)` is ordered without a synchronisation by the caller.
 *
 * Layouts: actions [n_envs][A] float32 row-major, obs [n_envs][O] float32, reward [n_envs] float32,
 * done [n_envs] uint8.  Generalised state [n_envs][S] float32 with
 * S = (floating base: pos3 quat4(xyzw) omega3 vel3) + q[ndof] + qd[ndof]   (SURVEY.md 8(d)).
 */
#ifndef ROBOSCHOOL_B200_H
#define ROBOSCHOOL_B200_H

#include <stdint.h>
#include "rs_model.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rs_world rs_world;
typedef struct rs_policy rs_policy;

const char* rs_last_error(void);

/* Number of CUDA devices visible (0 => every compute call fails). */
int rs_device_count(void);

/* Names that travel with a compiled model (b3GetBodyInfo / b3GetJointInfo strings, physics-bullet.cpp:133-251):
 * what the reference's env classes look parts and joints up by (gym_mujoco_xml_env.py:47-76). */
#define RS_NAME_LEN 48
typedef struct rs_model_names {
    char base_name[RS_NAME_LEN];                    /* root part ("world" for a jointed root)      */
    char link_names[RS_MAX_LINKS][RS_NAME_LEN];     /* body name; "dummy_<joint>" for importer-made links */
    char joint_names[RS_MAX_LINKS][RS_NAME_LEN];    /* "" for fixed joints                          */
    char geom_names[RS_MAX_GEOMS][RS_NAME_LEN];
} rs_model_names;

/* World::load_mjcf: MJCF file -> compiled model        physics-bullet.cpp:95-131 (b3LoadMJCFCommandInit with
 * URDF_USE_SELF_COLLISION | URDF_USE_SELF_COLLISION_EXCLUDE_ALL_PARENTS) + b3GetJointInfo :206-251.
 * Host-only (no CUDA device needed).  Fills the physics part of *out: topology, joint frames, mass / inertia / COM frames,
 * collision geoms, self-collision pairs, solver constants (physics-bullet.cpp:368-376); gravity and timestep are squeezed
 * through float32 like World(float gravity, float timestep) (python-binding.cpp:297).  The task constants (action gains,
 * feet, alive rule, costs: class attributes of gym_mujoco_walkers.py, not part of the MJCF) are left zero for the caller to
 * set; max_contacts = 16.  `names` may be NULL.  Same conventions and results as
 * roboschool_b200/mjcf.py with its default arguments. */
int rs_model_compile(const char* mjcf_path, double gravity, double timestep, int frame_skip, int has_floor,
                     rs_model_desc* out, rs_model_names* names);

/* World(gravity, timestep) + load_mjcf()            python-binding.cpp:297-301, 332-343;
 * physics-bullet.cpp:12-21, 95-131.  gravity/timestep/frame_skip come from the compiled model. */
int rs_world_create(const rs_model_desc* model, int n_envs, int device, rs_world** out);

/* World::~World / clean_everything                     physics-bullet.cpp:34-68 */
int rs_world_destroy(rs_world* w);

int rs_world_n_envs(const rs_world* w);
/* 0: runtime-topology step kernel; >0: step kernel specialised on a generated compile-time topology */
int rs_world_kernel_variant(const rs_world* w);
/* warps per CTA of the step kernel this world runs on (chosen from the batch size at creation, RS_WPC overrides).  Results are
 * bit-identical between worlds of the same model only when this number agrees (one compiled instantiation per value). */
int rs_world_cta_warps(const rs_world* w);
int rs_world_state_dim(const rs_world* w);
int rs_world_obs_dim(const rs_world* w);
int rs_world_act_dim(const rs_world* w);

/* env.reset() for a masked subset: joint U(-0.1,0.1), base pose, yaw U(+-pi/16)
 * (gym_mujoco_xml_env.py:39-78, gym_forward_walker.py:24-30, gym_mujoco_walkers.py:106-128;
 *  Joint::reset_current_position physics-bullet.cpp:568-577, World::robot_move :583-594).
 * mask_dev: uint8[n_envs] device pointer or NULL (= all).  obs_dev receives the reset observation
 * of the envs that were reset (may be NULL).  env_id0: global index of env 0 (multi-GPU sharding:
 * the RNG is keyed by global env id so results do not depend on the number of GPUs). */
int rs_world_reset(rs_world* w, uint32_t seed, uint32_t env_id0, const uint8_t* mask_dev, float* obs_dev, void* stream);

/* One env.step for every env: apply_action + World::step(frame_skip) + calc_state/reward/done
 * (gym_forward_walker.py:89-133; physics-bullet.cpp:358-412, 440-495, 547-566, 596-637).
 * auto_reset != 0: envs that finish (done or max_episode_steps) are reset in the same launch and
 * obs holds the first observation of the new episode (vector-env convention); done is still 1. */
int rs_world_step(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                  int auto_reset, void* stream);

/* rs_world_step that also accumulates rollout statistics into stats_dev[8] (float64, device):
 * [0] sum reward [1] episodes finished [2] sum episode length [3] non-finite obs [4] constraint rows
 * [5] contact points (last sub-step) [6] env-steps.  Warp-reduced, one atomicAdd per warp. */
int rs_world_step_stats(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                        int auto_reset, double* stats_dev, void* stream);

/* Profiling twin of rs_world_step (auto_reset on): also returns clock64() warp-cycle totals per phase in
 * cycles_host[16] = {kinematics, collide, aba, solve, integrate, epilogue+store, load+apply_action, total,
 * then inside the solve phase (static-topology kernels): enumerate rows, limit-row tasks, contact tasks, right-hand sides,
 * PGS iterations, 0, 0, 0}. */
int rs_world_step_profile(rs_world* w, const float* actions_dev, float* obs_dev, float* reward_dev, uint8_t* done_dev,
                          unsigned long long* cycles_host, void* stream);

/* Same call for callers holding HOST buffers (the reference's calling convention: numpy in, numpy
 * out): copies actions H2D, steps, copies obs/reward/done D2H on the world's stream and waits. */
int rs_world_step_host(rs_world* w, const float* actions_host, float* obs_host, float* reward_host,
                       uint8_t* done_host, int auto_reset);

/* policy + step for callers whose observations live in HOST memory (numpy): H2D obs_in -> policy forward on the
 * tensor cores -> env step with auto-reset -> D2H obs_out/reward/done; one stream synchronisation.
 * obs_in_host == NULL continues from the observations already on the device.  Page-locked host buffers (cudaHostAlloc /
 * cudaHostRegister, e.g. pinned torch tensors) are the source / target of the async copies directly; pageable ones go
 * through the world's pinned staging buffers.  With page-locked buffers and a static-topology kernel (>= 24576 envs) the envs are
 * processed as 6 CTA-aligned slices (the two outer ones half-size) on separate streams, so the copies of one slice overlap the kernels of the others; results are
 * bit-identical to the single-stream form (RS_E2E_CHUNKS=1..8 overrides the slice count). */
int rs_world_policy_step_host(rs_world* w, rs_policy* p, const float* obs_in_host, float* obs_out_host, float* reward_host,
                              uint8_t* done_host);

/* Physics only: n sub-steps with explicit joint torques tau_host[n_envs][ndof] (NULL = keep the
 * torques of the last call); no task logic.  Joint::set_motor_torque + World::bullet_step
 * (physics-bullet.cpp:497-508, 358-412).  A world with FlagrunHarder's cube (rs_model_desc.has_cube: World::load_urdf of
 * models_household/cube.urdf, physics-bullet.cpp:70-93) steps the cube too; Robot::set_pose_and_speed on it
 * (robot_move, :579-591) is rs_world_set_state on the last 13 state words. */
int rs_world_substeps(rs_world* w, int n_substeps, const float* tau_host);

/* Joint::set_servo_target(pos, kp, kd, maxforce) / Joint::set_target_speed(v, kd, maxforce) (physics-bullet.cpp:510-537;
 * python-binding.cpp:654-665): Bullet's btMultiBodyJointMotor of one joint dof -- a velocity-level row after the joint-limit rows,
 * target kp (target_pos - q) / dt + kd (target_vel - qd), impulse within -+ max_impulse per sub-step.  max_impulse <= 0 switches
 * the motor off (what Joint::set_motor_torque does on its first call).  env < 0: every env.  A world that has had a motor set
 * steps on the runtime-topology kernel (the static-topology kernels build no motor rows). */
int rs_world_set_joint_motor(rs_world* w, int env, int dof, float kp, float kd, float target_pos, float target_vel, float max_impulse);

/* State exchange with HOST buffers (tests, the single-env cpp_household shim, checkpointing). */
int rs_world_set_state(rs_world* w, const float* state_host, int clear_contacts);
int rs_world_get_state(rs_world* w, float* state_host);
/* potential float64[n], initial_z float32[n], feet_contact float32[n][n_feet], frame int32[n] */
int rs_world_set_episode(rs_world* w, const double* potential, const float* initial_z, const float* feet_contact, const int* frame);
int rs_world_get_episode(rs_world* w, double* potential, float* initial_z, float* feet_contact, int* frame);
/* Flagrun (gym_humanoid_flagrun.py:21-45): per-env flag position float64[n][2] (walk_target_x/y), env steps until it
 * moves int32[n] (flag_timeout) and RNG draws consumed by flag_reposition uint32[n].  NULL pointers are skipped. */
int rs_world_set_flags(rs_world* w, const double* target_xy, const int* timeout, const uint32_t* draws);
int rs_world_get_flags(rs_world* w, double* target_xy, int* timeout, uint32_t* draws);
/* FlagrunHarder (gym_humanoid_flagrun.py:73-169) per-env task state, float64[n][8]: task (0 walk, 1 stand up, 2 roll over),
 * on_ground_frame_counter, crawl_start_potential is set (0/1), crawl_start_potential, crawl_ignored_potential, and the
 * RepeatUnderlearnedTasks history of each task (bits 0-9 = last outcomes, newest in bit 0; bits 16+ = how many). */
int rs_world_set_task_state(rs_world* w, const double* state_host);
int rs_world_get_task_state(rs_world* w, double* state_host);

/* World::query_body_position (physics-bullet.cpp:440-495): per env, per body (base first, then
 * links): COM position3, quaternion4 (xyzw), linear velocity3  -> out_host[n_envs][1+L][10]. */
int rs_world_link_poses(rs_world* w, float* out_host);

/* World::bullet_contact_list (physics-bullet.cpp:596-637): live manifold points of one env in
 * solver order.  keys_host[max_pts]: manifold id (floor: geom index; self pair: n_geoms + pair
 * index); data_host[max_pts][10]: localA3 localB3 normal3 distance; *count: number of points. */
int rs_world_get_contacts(rs_world* w, int env, int* keys_host, float* data_host, int max_pts, int* count);
/* Overwrite the persistent contact manifolds of every env (test / checkpoint hook; the inverse of rs_world_get_contacts):
 * counts_host[n_envs], keys_host[n_envs][max_pts], data_host[n_envs][max_pts][10], points in solver order. */
int rs_world_set_contacts(rs_world* w, const int* counts_host, const int* keys_host, const float* data_host, int max_pts);
/* rows / contact points of the last solved sub-step, int32[n_envs] each (host) */
int rs_world_get_counts(rs_world* w, int* rows_host, int* contacts_host);
/* last step's reward terms [n_envs][5] (alive, progress, electricity, joints_at_limit, feet_collision) */
int rs_world_get_rewards5(rs_world* w, float* out_host);

/* Device pointer to the SoA state block [S][n_envs] (zero-copy hand-off to torch). */
int rs_world_state_dev(rs_world* w, float** state_dev);

/* agent_zoo v1 policy: relu(relu(x W1 + b1) W2 + b2) W3 + b3
 * (agent_zoo/RoboschoolHumanoid_v1_2017jul.py:21-34,63-68).  Weights: host float32, row-major
 * W1[in][h1] b1[h1] W2[h1][h2] b2[h2] W3[h2][out] b3[out]. */
int rs_policy_create(int in_dim, int h1, int h2, int out_dim, const float* w1, const float* b1, const float* w2,
                     const float* b2, const float* w3, const float* b3, int device, rs_policy** out);
int rs_policy_destroy(rs_policy* p);
/* act_dev[n][out] = policy(obs_dev[n][in]) on the tensor cores (tcgen05.mma, fp16 operands, fp32 accumulation in TMEM) */
int rs_policy_forward(rs_policy* p, const float* obs_dev, float* act_dev, int n, void* stream);

/* Fused rollout: n_steps x { act = policy(obs) ; step(auto_reset) } with no host round trip; the
 * rollout statistics (sum reward, episodes finished, sum episode length, non-finite count, contact
 * rows) are accumulated on the device into stats_dev[8] (float64).  policy may be NULL when
 * actions are uniform random (counter-based RNG).  Auto-resets and random actions use the seed / env_id0 given to the
 * last rs_world_reset / rs_world_rollout_reset, like every other stepping entry point. */
int rs_world_rollout(rs_world* w, rs_policy* p, int n_steps, double* stats_dev, void* stream);
/* reset every env and write the first observations into the world-owned rollout obs buffer */
int rs_world_rollout_reset(rs_world* w, uint32_t seed, uint32_t env_id0, void* stream);
/* device pointers of the world-owned rollout buffers: obs [n][O], act [n][A], reward [n], done [n] (uint8) */
int rs_world_rollout_buffers(rs_world* w, float** obs_dev, float** act_dev, float** rew_dev, uint8_t** done_dev);
/* Rollout recorder for trainers: n_steps policy+env steps written into the caller's time-major DEVICE
 * tensors obs_T [n_steps+1][n][O], act_T [n_steps][n][A], rew_T [n_steps][n], done_T [n_steps][n] (uint8); obs_T[0] is
 * filled from obs0_dev [n][O], or, if that is NULL, with the world's own current observations (after
 * rs_world_rollout_reset or a previous rollout).  Finished envs are
 * reset on the device (done=1, next obs = first obs of the new episode).  policy == NULL replays the actions in act_T. */
int rs_world_rollout_record(rs_world* w, rs_policy* policy, int n_steps, const float* obs0_dev, float* obs_T, float* act_T,
                            float* rew_T, uint8_t* done_T, double* stats_dev, void* stream);


#ifdef __cplusplus
}
#endif
#endif
