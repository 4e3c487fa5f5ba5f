/*
 * rt_env.h — C ABI of the B200-native batched radiotherapy environment step.
 *
 * The reference (rmaguado/ppo-radiotherapy) is pure Python and has no FFI layer
 * (SURVEY.md §8b); this header is the boundary a maintainer binds with ctypes to
 * replace the reference functions cited at each entry point.  Conventions:
 *   - plain C symbols, plain pointers and sizes, no torch / Python types;
 *   - every call returns 0 on success and a negative rt_status on failure, with a
 *     thread-local message available from rt_last_error();
 *   - pointers named *_dev are DEVICE pointers owned by the caller (e.g. PyTorch
 *     allocations); calls taking a `stream` are asynchronous and stream-ordered and
 *     never synchronise the host; `stream` is a cudaStream_t passed as void*;
 *   - pointers named *_host are HOST pointers; the *_host calls run the same kernels on the
 *     handle's own stream and return after it has drained.  Page-locked buffers are mapped
 *     into the kernel (zero-copy), pageable ones are staged through the handle's pinned
 *     buffers.  A *_host call that follows device-pointer calls on the same handle waits for
 *     them first.  The pinned-or-not answer is cached per address: do not free a pinned
 *     buffer and pass a pageable array that reuses its address while a handle lives;
 *   - one handle drives one device and is not thread-safe.
 *
 * Array layouts are C-order.  G = grid = (67, 43, 70) for the bundled phantom,
 * V = G0*G1*G2 voxels, linear voxel index = (i*G1 + j)*G2 + k (numpy C order of
 * the reference's volumes, environment.py:29-30).
 */
#ifndef RT_ENV_H
#define RT_ENV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define RT_API __attribute__((visibility("default")))
#else
#define RT_API
#endif

#define RT_ABI_VERSION 2

typedef enum {
    RT_OK = 0,
    RT_ERR_INVALID = -1,   /* bad argument */
    RT_ERR_CUDA = -2,      /* CUDA runtime error (message has the detail) */
    RT_ERR_NOMEM = -3,
    RT_ERR_STATE = -4      /* call not valid in the handle's current state */
} rt_status;

/* Environment constants (environment.py:16-26). */
#define RT_ACTION_SIZE 6
#define RT_OBS_SIZE 9
#define RT_MAX_TIME_STEPS 100
#define RT_INFO_SIZE 16
/* Upper bound on distinct voxels of one beam for the bundled grid: 4 * (max(G) + 1) = 284 splat
 * writes (draw_line.py:68-96), rounded up.  Other grids need cap >= 4 * (max(G) + 1). */
#define RT_BEAM_CAP 288

/* rt_create flags */
#define RT_FLAG_DENSE 1u         /* full-volume dose update + from-scratch reductions every step
                                    (the reference's own dataflow, environment.py:107-110,164-191) */
#define RT_FLAG_RECORD_BEAMS 2u  /* keep the per-episode beam list (environment.py:110 self.beams) */

/* Columns of the optional info block written by rt_step (float64 [N][RT_INFO_SIZE]);
 * the reference's info dict, environment.py:222-241, plus episode statistics
 * (gymnasium RecordEpisodeStatistics, train.py:36). */
enum {
    RT_INFO_REWARD_TOTAL = 0,   /* reward_components.total  */
    RT_INFO_REWARD_TUMOUR = 1,  /* reward_components.tumour */
    RT_INFO_REWARD_LUNG = 2,    /* reward_components.lung   */
    RT_INFO_REWARD_DISTANCE = 3,/* reward_components.distance_to_tumour */
    RT_INFO_DOSE_TUMOUR = 4,    /* doses.tumour */
    RT_INFO_DOSE_LUNG = 5,      /* doses.lung   */
    RT_INFO_OVERSHOOT_T0 = 6,   /* overshoot.translation[0..2] */
    RT_INFO_OVERSHOOT_R = 9,    /* overshoot.rotation */
    RT_INFO_EPISODE_RETURN = 10,/* episode.r (valid where terminated) */
    RT_INFO_EPISODE_LENGTH = 11,/* episode.l */
    RT_INFO_LUNG_COUNT = 12,    /* voxels of lungs\tumour with dose > 0.2 (environment.py:176-177) */
    RT_INFO_STEPPED = 13,       /* 1 if this call advanced the env, 0 if it was the autoreset call */
    RT_INFO_TUMOUR_ID = 14,
    RT_INFO_T = 15
};

/* Phantom description, HOST pointers, copied to the device by rt_create.
 * Built from the reference's data/lungs.npy and data/tumours/*.npy
 * (environment.py:28-29,90-97) by tools/pack_phantom.py. */
typedef struct {
    int32_t grid[3];
    const uint32_t *lungs_bits;    /* ceil(V/32) words, bit v = lungs.flat[v] */
    int32_t n_tumours;
    const int32_t *vox_offsets;    /* [n_tumours + 1] */
    const int32_t *vox;            /* ascending linear voxel indices per tumour */
    const double *centroid;        /* [n_tumours][3], environment.py:145-148 */
    const float *tumour_sum;       /* [n_tumours], np.sum(tumours), environment.py:167 */
    const float *lung_mask_sum;    /* [n_tumours], np.sum(lungs*(1-tumours)), environment.py:178 */
} rt_phantom_desc;

typedef struct rt_env rt_env;

RT_API int rt_abi_version(void);
RT_API const char *rt_last_error(void);

/* ---- lifetime ----------------------------------------------------------------- */
/* N resident episodes on CUDA device `device` (replaces constructing N
 * RadiotherapyEnv objects inside SyncVectorEnv, train.py:93-95).  Episodes start
 * un-reset: call rt_reset before rt_step. */
RT_API int rt_create(rt_env **out, int device, int n_envs, uint32_t flags, const rt_phantom_desc *phantom);
RT_API int rt_destroy(rt_env *env);
RT_API int rt_num_envs(const rt_env *env);
RT_API int64_t rt_device_bytes(const rt_env *env);

/* ---- tumour choice (environment.py:90 np.random.choice) --------------------------- */
/* Counter-based device RNG: episode e of env i uses tumour hash(seed, i, e) mod n_tumours. */
RT_API int rt_seed(rt_env *env, uint64_t seed);
/* Explicit schedule: episode e of env i uses ids_host[min(e, n_episodes-1)][i]; NULL returns to the RNG. */
RT_API int rt_set_tumour_schedule(rt_env *env, const int32_t *ids_host, int n_episodes);

/* ---- reset / step (environment.py:77-105, 193-243; SyncVectorEnv NEXT_STEP autoreset) ---- */
/* Reset every env whose mask byte is non-zero (mask_dev NULL = all) and restart its
 * episode counter at 0; obs_dev float32 [N][9] receives the observation of EVERY env. */
RT_API int rt_reset(rt_env *env, const uint8_t *mask_dev, float *obs_dev, void *stream);

/* One vector step.  actions_dev float32 [N][6].  Outputs (any may be NULL except obs_dev):
 *   obs_dev float32 [N][9]      get_vector_observation, environment.py:259-268
 *   reward_dev float64 [N]      environment.py:218     reward_f32_dev float32 [N] (same, rounded)
 *   terminated_dev uint8 [N]    environment.py:220     truncated_dev uint8 [N] (always 0, :243)
 *   info_dev float64 [N][RT_INFO_SIZE]
 * An env that terminated on the previous call ignores its action, resets, and reports
 * reward 0 / terminated 0 (gymnasium 1.0.0 AutoresetMode.NEXT_STEP, train.py:151). */
RT_API int rt_step(rt_env *env, const float *actions_dev, float *obs_dev, double *reward_dev,
                   float *reward_f32_dev, uint8_t *terminated_dev, uint8_t *truncated_dev,
                   double *info_dev, void *stream);

/* Host-buffer forms of the same calls (the reference's seam: train.py:151-158 passes
 * numpy arrays).  Synchronous. */
RT_API int rt_reset_host(rt_env *env, const uint8_t *mask_host, float *obs_host);
RT_API int rt_step_host(rt_env *env, const float *actions_host, float *obs_host, double *reward_host,
                        uint8_t *terminated_host, uint8_t *truncated_host, double *info_host);

/* ---- state access ---------------------------------------------------------------- */
/* pose_dev float64 [N][6] = (beam_position, beam_direction), environment.py:48-49. */
RT_API int rt_get_pose(rt_env *env, double *pose_dev, void *stream);
RT_API int rt_set_pose(rt_env *env, const double *pose_dev, void *stream);
/* counters_dev int32 [N][6] = (t, tumour id, lung voxels above threshold, episode index,
 * needs-reset flag, beams recorded) — environment.py:47 self.t and friends. */
RT_API int rt_get_counters(rt_env *env, int32_t *counters_dev, void *stream);
/* Dense float32 dose volume of one env, [V] (environment.py:42 self.dose). */
RT_API int rt_get_dose(rt_env *env, int env_index, float *dose_dev, void *stream);
/* Voxel observation, float32 [N][4][V] = clip(stack[lungs, tumours, dose, view], 0, 1)
 * (get_volumes, environment.py:245-257); envs [first, first+count). */
RT_API int rt_assemble_volumes(rt_env *env, int first, int count, float *obs_dev, void *stream);
/* Compressed voxel-observation records for rollout storage (train.py:110-112 keeps num_steps x num_envs float32
 * observations: 3.2 MB each).  A record is the dose volume as bfloat16 [stride] (stride =
 * rt_observation_record_stride(), V rounded up to 32), the pose float64 [6] and the tumour id; lungs, tumour and
 * beam-view planes are regenerated from them.  rt_pack_observations writes the records of envs
 * [first, first+count) to slots [slot0, slot0+count) of caller-owned arrays; rt_render_observations rebuilds
 * float32 [count][4][V] for records index_dev[0..count) (or 0..count-1 when index_dev is NULL).  The dose plane of
 * a rendered observation is the bfloat16-rounded dose, which is what a bf16 convolution reads anyway; the other
 * three planes are identical to rt_assemble_volumes. */
RT_API int rt_observation_record_stride(const rt_env *env);
RT_API int rt_pack_observations(rt_env *env, int first, int count, int64_t slot0, void *dose_bf16_dev, double *pose_dev,
                                int32_t *tumour_id_dev, void *stream);
RT_API int rt_render_observations(rt_env *env, const void *dose_bf16_dev, const double *pose_dev,
                                  const int32_t *tumour_id_dev, const int64_t *index_dev, int count, float *obs_dev,
                                  void *stream);
/* Recorded beams of one env (needs RT_FLAG_RECORD_BEAMS): float64 [100][6], returns count via n_dev. */
RT_API int rt_get_beams(rt_env *env, int env_index, double *beams_dev, int32_t *n_dev, void *stream);

/* ---- stateless geometry entry points (parity surface) ------------------------------- */
/* draw_line.py:4 beam_voxels for m rays.  pos_dev/dir_dev float64 [m][3].  Per ray the
 * distinct voxels hit and their summed float32 weights: idx_dev int32 [m][cap], w_dev
 * float32 [m][cap], count_dev int32 [m]; count -1 flags the ValueError of draw_line.py:23-24
 * ("Direction vector magnitude is too small.").  cap >= 4 * (max(grid) + 1). */
RT_API int rt_beam_voxels(const int32_t grid[3], const double *pos_dev, const double *dir_dev, int m, int cap,
                          int32_t *idx_dev, float *w_dev, int32_t *count_dev, void *stream);
/* Same, written as the reference returns it: dense float32 [m][V] (zero-filled here). */
RT_API int rt_beam_voxels_dense(const int32_t grid[3], const double *pos_dev, const double *dir_dev, int m,
                                float *out_dev, int32_t *status_dev, void *stream);
/* environment.py:112-143 + transforms.py:7-69 for m independent poses: actions float32 [m][6]
 * -> new position / direction float64 [m][3], translation overshoot [m][3], rotation overshoot [m]. */
RT_API int rt_pose_update(const int32_t grid[3], const double *pos_dev, const double *dir_dev,
                          const float *actions_dev, int m, double *pos_out_dev, double *dir_out_dev,
                          double *overshoot_t_dev, double *overshoot_r_dev, void *stream);

/* transforms.py:7 apply_rotation(initial_direction, rotation_vector, min_angle) for m vectors:
 * float64 [m][3] directions and rotation vectors -> new directions [m][3], overshoot [m]. */
RT_API int rt_apply_rotation(const double *dir_dev, const double *rotvec_dev, int m, double min_angle,
                             double *dir_out_dev, double *overshoot_dev, void *stream);
/* transforms.py:62 apply_translation(position, translation_vector, bounds): clip(p + t, 0, bounds)
 * and |p + t - clipped| for m positions; bounds is a HOST float64[3]. */
RT_API int rt_apply_translation(const double *pos_dev, const double *translation_dev, int m,
                                const double bounds[3], double *pos_out_dev, double *overshoot_dev,
                                void *stream);

/* ---- GAE (train.py:164-181) -------------------------------------------------------- */
/* float32 [T][N] rewards/values/dones, [N] next_value/next_done -> advantages, returns [T][N]. */
RT_API int rt_gae(const float *rewards_dev, const float *values_dev, const float *dones_dev,
                  const float *next_value_dev, const float *next_done_dev, int T, int N,
                  double gamma, double gae_lambda, float *advantages_dev, float *returns_dev, void *stream);

/* ---- rollout side of the PPO loop (train.py:139-161, networks.py:107-147), SURVEY.md 8f-1 ------------------ */
/* Device pointers to the parameters of the reference's MLP agent `PPO` (networks.py:107-130; state-dict keys
 * critic.{0,2,4}.{weight,bias}, actor_mean.{0,2,4}.{weight,bias}, actor_logstd), torch Linear layout weight [out][in].
 * Supported: n_obs <= 16, hidden == 64 (feature_dim of configs/*.yaml.template), n_act <= 6. */
typedef struct rt_mlp_params {
    const float *critic_w0, *critic_b0, *critic_w1, *critic_b1, *critic_w2, *critic_b2;
    const float *actor_w0, *actor_b0, *actor_w1, *actor_b1, *actor_w2, *actor_b2;
    const float *actor_logstd;
    int32_t n_obs, hidden, n_act;
} rt_mlp_params;
/* train.py:139-149 in one kernel: row t = counters_dev[0] of the rollout buffers obs [T][n][n_obs], dones [T][n] receives
 * obs_dev / next_done_dev; `agent.get_action_and_value(next_obs)` (networks.py:132-147, no_grad) is evaluated in float32;
 * action = mean + exp(logstd) * N(0,1) from Philox4x32-10 keyed by (seed, env, counters_dev[1]); values [T][n],
 * actions [T][n][n_act], logprobs [T][n] receive row t; action_out_dev [n][n_act] is the input of rt_step.  Any rollout
 * buffer may be NULL.  `rows` is the row count T of the rollout buffers: when counters_dev[0] is outside [0, rows) the
 * kernel still produces the actions but stores no row (a caller that steps past the end of its buffers does not write
 * past them).  `p` is a HOST struct of device pointers. */
RT_API int rt_ppo_act(const rt_mlp_params *p, const float *obs_dev, const float *next_done_dev, int n, uint64_t seed,
                      const int64_t *counters_dev, int rows, float *obs_buf_dev, float *dones_buf_dev, float *values_buf_dev,
                      float *actions_buf_dev, float *logprobs_buf_dev, float *action_out_dev, void *stream);
/* train.py:153-161 after rt_step: rewards [T][n] row t = counters_dev[0] <- reward_f32_dev, next_done_dev [n] <-
 * terminated | truncated, and for the envs that terminated the episode statistics train.py:42-66 logs
 * (episode_stats_dev float64 [7]: finished, sum of episode returns, lengths, last-step tumour / lung / distance / total
 * reward; accumulated, the caller zeroes it) from info_dev [n][RT_INFO_SIZE].  truncated, info, rewards, stats may be NULL. */
RT_API int rt_ppo_record(const float *reward_f32_dev, const uint8_t *terminated_dev, const uint8_t *truncated_dev,
                         const double *info_dev, int n, const int64_t *counters_dev, int rows, float *rewards_buf_dev,
                         float *next_done_dev, double *episode_stats_dev, void *stream);

/* The whole rollout of train.py:138-161 in ONE launch: n_steps times { obs / done -> row, policy forward + sample +
 * log-prob -> value / action / log-prob rows, environment step, reward -> row, done -> next_done, episode statistics },
 * every block keeping its envs for all steps (no grid-wide barrier between steps).  Rows row0 .. row0 + n_steps - 1 of
 * the rollout buffers (obs [rows][n][9], dones / values / logprobs / rewards [rows][n], actions [rows][n][6]) are written;
 * the call fails if they do not fit in `rows`.  next_obs_dev [n][9] / next_done_dev [n] hold the observation / done flag
 * before the first step on entry (e.g. the obs_dev of rt_reset) and after the last step on return.  The random stream
 * is the one of rt_ppo_act: step t uses Philox key (seed, env, rng_step0 + t), so the rows equal n_steps x (rt_ppo_act,
 * rt_step, rt_ppo_record) bit for bit.  The agent must be the reference's MLP for this env (n_obs 9, hidden 64, n_act 6).
 * Sparse-mode handles only. */
RT_API int rt_rollout(rt_env *env, const rt_mlp_params *p, int n_steps, int64_t row0, int rows, uint64_t seed, int64_t rng_step0,
                      float *obs_buf_dev, float *dones_buf_dev, float *values_buf_dev, float *actions_buf_dev,
                      float *logprobs_buf_dev, float *rewards_buf_dev, float *next_obs_dev, float *next_done_dev,
                      double *episode_stats_dev, void *stream);

/* ---- FeaturesExtractor3D, first block (networks.py:15-24) ------------------------------------------ */
/* Conv3d(4->16, k=3) + bias + ReLU + MaxPool3d(2, 2, padding=((D-2)%2, (H-2)%2, (W-2)%2)) fused in one tensor-core
 * kernel (tcgen05, accumulators in tensor memory).  x_dev float32 [n][4][D][H][W] (the voxel observation),
 * weight_dev float32 [16][4][3][3][3], bias_dev float32 [16] -> out_dev bfloat16 [n][16][Pd][Ph][Pw] (NCDHW),
 * P = (dim - 2 + pad - 2)/2 + 1.  scratch_dev: 16,384 bytes of device memory for the repacked weights.  W must be even (no pool padding on the last axis);
 * RT_ERR_INVALID is returned for shapes the kernel does not cover so that the caller can use its own path. */
RT_API int rt_conv1_relu_pool(const float *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H,
                              int W, void *out_dev, void *scratch_dev, void *stream);
/* Same block with the output in the grouped channels-last layout the second block reads:
 * out_dev bfloat16 [n][2 groups][Pd][Ph*Pw][8 channels] (channel = group*8 + c). */
RT_API int rt_conv1_relu_pool_grouped(const float *x_dev, const float *weight_dev, const float *bias_dev, int n, int D,
                                      int H, int W, void *out_dev, void *scratch_dev, void *stream);
/* The first block computed straight from the state of envs [first, first+count): the voxel observation is generated
 * inside the kernel and never written to HBM.  Output as rt_conv1_relu_pool_grouped, bit-identical to
 * rt_assemble_volumes followed by it.  Sparse-mode handles only. */
RT_API int rt_conv1_from_env(rt_env *env, int first, int count, const float *weight_dev, const float *bias_dev,
                             void *out_dev, void *scratch_dev, void *stream);
/* Second block (networks.py:25-27): Conv3d(16->16, k=3, groups=2) + bias + ReLU + MaxPool3d(2, 2) on tcgen05.
 * x_dev bfloat16 [n][2][D][H*W][8] as written by rt_conv1_relu_pool_grouped, weight_dev float32 [16][8][3][3][3],
 * bias_dev float32 [16] -> out_dev bfloat16 [n][16][(D-2)/2][(H-2)/2][(W-2)/2] (NCDHW).  scratch_dev: 65,536 bytes.
 * W must be even; RT_ERR_INVALID for shapes the kernel does not cover. */
RT_API int rt_conv2_relu_pool(const void *x_dev, const float *weight_dev, const float *bias_dev, int n, int D, int H,
                              int W, void *out_dev, void *scratch_dev, void *stream);
/* Tail (networks.py:28-45): Conv3d(16->16, k=3, groups=4) + ReLU + MaxPool3d(2, 2) + Flatten + Linear + ReLU, one
 * kernel.  x_dev bfloat16 [n][16][D][H][W] (the second block's output), conv_w_dev float32 [16][4][3][3][3],
 * conv_b_dev [16], lin_w_dev float32 [F][16*Pd*Ph*Pw] (P = (dim-2)/2; flatten order c, d, h, w), lin_b_dev [F], F <= 256
 * -> out_dev float32 [n][F] (the features FeaturesExtractor3D.forward returns). */
RT_API int rt_c3d_tail(const void *x_dev, const float *conv_w_dev, const float *conv_b_dev, const float *lin_w_dev,
                       const float *lin_b_dev, int n, int D, int H, int W, int F, float *out_dev, void *stream);

/* ---- instrumentation ----------------------------------------------------------------- */
/* Number of kernels this library has launched since load (for bench.py's gpu_launches). */
RT_API int64_t rt_launch_count(void);
/* Developer aid: when stamps_dev (int64 [N][12], device) is non-NULL the step kernel records clock64()
 * at its stage boundaries per env (scalar warp: 0 start, 8 state loaded and translated, 9 pose updated,
 * 10 beam set up, 1 walk done, 11 past barrier 2, 7 end; env warp: 2 tumour entry + distance done, 3 past
 * barrier 1, 4 first pass's cell loads issued, 6 all passes stored).  Sparse-mode handles with more than
 * 7 envs per SM only.  NULL (the default) switches it off. */
RT_API int rt_set_stage_clock(rt_env *env, long long *stamps_dev);
/* Measurement aid: consecutive rt_step launches overlap through programmatic dependent launch (the next
 * launch's blocks are scheduled while the previous one drains).  enabled = 0 switches that off for A/B timing. */
RT_API int rt_set_pdl(rt_env *env, int enabled);

#ifdef __cplusplus
}
#endif
#endif /* RT_ENV_H */
