/*
 * d2d_b200.h -- C ABI of the B200-native D2D-PPO hot path (libd2d_b200.so).
 *
 * The reference (benrobaglia/D2D-PPO) is pure Python and defines no FFI; its
 * boundary for this path is the Python env/agent API.  Each entry point below
 * names the reference function (file:line under /root/reference) whose work it
 * replaces.  The Python classes in d2d-ppo_b200/ keep the reference signatures
 * and call these through ctypes (INTEGRATION.md shows the binding).
 *
 * Conventions
 *   - extern "C", plain C types; no torch / C++ types cross the boundary.
 *   - every function returns int: 0 = ok, negative = D2D_ERR_*; the message is
 *     in d2d_last_error() (thread-local).  Nothing throws.
 *   - all data pointers are DEVICE pointers borrowed for the duration of the
 *     call (PyTorch owns them) unless a parameter is documented as "host".
 *   - calls are asynchronous on the given cudaStream_t (passed as void*); no
 *     hidden synchronisation.
 *   - a handle is not thread-safe; distinct handles are independent.
 *
 * Device layout ("env-minor" structure of arrays): every per-(env, device)
 * quantity is stored as X[row][env] with the env index fastest, so that a warp
 * of 32 consecutive envs reads/writes 32 consecutive elements.  B = n_envs.
 */
#ifndef D2D_B200_H
#define D2D_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define D2D_ABI_VERSION 1

#define D2D_OK 0
#define D2D_ERR_INVALID (-1)   /* bad argument / unsupported shape */
#define D2D_ERR_CUDA (-2)      /* a CUDA runtime call failed */
#define D2D_ERR_STATE (-3)     /* call sequence error (e.g. step before reset) */

/* env kinds */
#define D2D_ENV_COMBINATORIAL 0      /* envs/combinatorial_env.py:4  CombinatorialEnv    */
#define D2D_ENV_SINGLE_CHANNEL 1     /* envs/env.py:4                D2DEnv              */
#define D2D_ENV_CHANNEL_SELECTION 2  /* envs/channel_selection_env.py:4 ChannelSelectionEnv */

/* random streams */
#define D2D_RNG_PHILOX 0  /* counter-based Philox4x32-10 keyed by (seed, env, t, device) */
#define D2D_RNG_REPLAY 1  /* pre-drawn arrival / switch streams (parity with the reference) */

/* arrival distributions (combinatorial_env.py:180 poisson, :185/:195 binomial) */
#define D2D_ARRIVAL_POISSON 0
#define D2D_ARRIVAL_BERNOULLI 1

#define D2D_MAX_AGENTS 64
#define D2D_MAX_CHANNELS 32
#define D2D_MAX_DEADLINE 32
#define D2D_POISSON_KMAX 16

const char* d2d_last_error(void);
int d2d_abi_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches) */
uint64_t d2d_launch_count(void);

/* ------------------------------------------------------------------------------------------
 * Environments.  Replaces CombinatorialEnv / D2DEnv / ChannelSelectionEnv __init__/reset/step.
 * All pointers in the config are HOST pointers, copied by d2d_env_create.
 * ------------------------------------------------------------------------------------------ */
typedef struct d2d_env_config {
  int32_t kind;              /* D2D_ENV_*                                                   */
  int32_t n_envs;            /* B lockstep env instances on this GPU                        */
  int32_t n_agents;          /* N <= D2D_MAX_AGENTS                                         */
  int32_t n_channels;        /* C (1 for D2D_ENV_SINGLE_CHANNEL); C+1 <= 32 for selection   */
  int32_t episode_length;    /* T: done = (timestep >= T)   combinatorial_env.py:233        */
  int32_t homogeneous_size;  /* combinatorial_env.py:47-53: pad buffer obs to max deadline  */
  int32_t rng_mode;          /* D2D_RNG_*                                                   */
  int32_t reserved0;
  uint64_t seed;             /* Philox key                                                  */
  uint64_t env_offset;       /* global index of local env 0 (multi-GPU sharding)            */
  const int32_t* deadlines;        /* [N], each in 1..D2D_MAX_DEADLINE                      */
  const int32_t* arrival_kind;     /* [N] D2D_ARRIVAL_*                                     */
  const uint64_t* arrival_active;  /* [T+1] bit k: device k draws an arrival at timestep t
                                      (t = 0 is reset); encodes traffic_model / period / offsets */
  const uint32_t* poisson_cdf;     /* [N][D2D_POISSON_KMAX]: floor(CDF(m) 2^32), clamped    */
  const uint64_t* bernoulli_thr;   /* [N]: round(arrival_prob 2^32) in 0..2^32              */
  const uint32_t* switch_thr;      /* round(p 65536) in 0..65536; combinatorial [N][C],
                                      single-channel [N], selection [C+1]                   */
  const int32_t* nbr_offset;       /* single-channel only: CSR neighbourhoods, [N+1]; NULL = self */
  const int32_t* nbr_index;        /* single-channel only: [nbr_offset[N]]                  */
} d2d_env_config;

typedef struct d2d_env d2d_env;

int d2d_env_create(const d2d_env_config* cfg, d2d_env** out);
int d2d_env_destroy(d2d_env* env);

/* layout queries (rows of the env-minor obs / state matrices) */
int d2d_env_obs_rows(const d2d_env* env);               /* sum_k obs_dim[k]                 */
int d2d_env_obs_offset(const d2d_env* env, int agent);  /* first row of device k            */
int d2d_env_obs_dim(const d2d_env* env, int agent);     /* observation_space[k].shape[0]    */
int d2d_env_state_rows(const d2d_env* env);             /* state_space.shape[0]             */
int d2d_env_timestep(const d2d_env* env);

/* Replay streams (device, borrowed until the next set_replay/destroy):
 *   arrivals  u8 [t_len][N][B]           value device k would draw at timestep t (0 = reset)
 *   switches  combinatorial: mask [t_len][N][B], element = 1/2/4 bytes for C <= 8/16/32, bit c = flip
 *             single-channel: u8 [t_len][N][B] (bit 0); selection: u32 [t_len][B] (bit c, c = 0..C)
 *   index 0 of `switches` is unused (reset draws no switch).                                  */
int d2d_env_set_replay(d2d_env* env, const uint8_t* arrivals, const void* switches, int t_len);

/* reset(): combinatorial_env.py:61-114, env.py:51-101, channel_selection_env.py:49-98.
 *   obs   f32 [obs_rows][B]   (required)
 *   state f32 [state_rows][B] (may be NULL)                                                   */
int d2d_env_reset(d2d_env* env, float* obs, float* state, void* stream);

/* step(actions): combinatorial_env.py:127-242, env.py:118-217, channel_selection_env.py:116-214.
 *   actions  combinatorial: channel bitmask [N][B] (1/2/4 bytes per element, bit c = transmit on c)
 *            single-channel: u8 [N][B] 0/1;  selection: u8 [N][B] channel id 0..C (0 = idle)
 *   obs      f32 [obs_rows][B]       next observations (may be NULL: not emitted)
 *   state    f32 [state_rows][B]     next global state (may be NULL)
 *   reward   i32 [B]   combinatorial/selection: number of successful devices; single-channel: ack
 *   done     u8  [B]   (may be NULL)
 *   ack      combinatorial: i8 [C][B]; single-channel: NULL; selection: f32 [C+1][B] (may be NULL)  */
int d2d_env_step(d2d_env* env, const void* actions, float* obs, float* state, int32_t* reward,
                 uint8_t* done, void* ack, void* stream);

/* step with the fused random-access policy: algorithms/baselines.py:181-183
 * (CombinatorialRandomAccess.act = Bernoulli(tp) per (device, channel)); the action bits come from
 * the Philox policy stream inside the step kernel.  actions_out (same layout as `actions`) may be NULL. */
int d2d_env_step_random_access(d2d_env* env, double transmission_prob, void* actions_out, float* obs,
                               float* state, int32_t* reward, uint8_t* done, void* ack, void* stream);

/* pack reference-layout actions u8 [B][N][C] (0/1) into the bitmask layout above */
int d2d_pack_actions(const uint8_t* actions_bnc, void* packed, int n_envs, int n_agents, int n_channels,
                     void* stream);

/* Counters and raw state, env-minor: buffers u8 [N][Dmax][B]... exported as
 *   buffers  u8  [N][B][rec]   rec = d2d_env_record_bytes() (8/16/32), byte d = packets with d slots left
 *   channel  combinatorial: mask [N][B]; single-channel u8 [N][B]; selection u32 [B]
 *   discarded / received  u32 [N][B]     (combinatorial_env.py:91-92,174,181)
 *   stats    u32 [2][B]   single-channel: channel_errors, n_collisions (env.py:147,150);
 *                         selection: selected_channel_qualities, number_selected_channel (:132-133)
 * Any pointer may be NULL.  import is the inverse (used by tests to force a state). */
int d2d_env_record_bytes(const d2d_env* env);
int d2d_env_mask_bytes(const d2d_env* env);
int d2d_env_export_state(const d2d_env* env, uint8_t* buffers, void* channel, uint32_t* discarded,
                         uint32_t* received, uint32_t* stats, void* stream);
int d2d_env_import_state(d2d_env* env, const uint8_t* buffers, const void* channel, const uint32_t* discarded,
                         const uint32_t* received, const uint32_t* stats, int timestep, void* stream);

/* compute_urllc / compute_jains / compute_channel_score per env
 * (combinatorial_env.py:245-264): f64 [B] each, any may be NULL. */
int d2d_env_scores(const d2d_env* env, double* urllc, double* jains, double* channel_score, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* D2D_B200_H */
