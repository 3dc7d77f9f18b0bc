/*
 * qrmsa_b200.h -- C ABI of the B200-native batched QRMSA environment step.
 *
 * The reference (LEA-UFPA/optical-networking-gym) has no C ABI: its native boundary for this path
 * is the Cython extension module `optical_networking_gym.envs.qrmsa`, reached through Python
 * attribute access on a `QRMSAEnv` instance.  Each entry point below names the reference
 * interface it stands in for (file:line under /root/reference).  The Python host layer
 * (optical_networking_gym_b200/) binds these with ctypes and re-creates the QRMSAEnv
 * reset/step/action-mask surface on top; INTEGRATION.md shows the reference-side stub.
 *
 * Conventions: plain C types only; every function returns a qrmsa_status (0 = OK) unless noted;
 * `stream` is a cudaStream_t passed as void* (NULL = default stream); pointers named d_* are
 * device pointers on the context's device, h_* are host pointers (pinned or pageable).  One
 * context per device; contexts share no global state, so eight can coexist in one process.
 * There is NO CPU fallback: without a CUDA device qrmsa_create fails with QRMSA_ERR_NO_DEVICE.
 */
#ifndef QRMSA_B200_H
#define QRMSA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct qrmsa_ctx qrmsa_ctx;
typedef struct qrmsa_tracegen qrmsa_tracegen;

typedef enum qrmsa_status {
    QRMSA_OK = 0,
    QRMSA_ERR_ARG = 1,         /* bad argument                                             */
    QRMSA_ERR_CUDA = 2,        /* CUDA runtime error; see qrmsa_last_error                 */
    QRMSA_ERR_UNSUPPORTED = 3, /* configuration outside what the kernels implement         */
    QRMSA_ERR_STATE = 4,       /* call order (e.g. step before a trace is loaded)          */
    QRMSA_ERR_NO_DEVICE = 5,   /* no usable CUDA device: the product has no CPU path       */
    QRMSA_ERR_ENV = 6          /* an environment raised: the reference would have thrown   */
} qrmsa_status;

/*
 * Static description of one network + traffic classes, exported from the reference `topology`
 * graph (reference topology.pyx:244-369: graph["ksp"], ["modulations"], ["node_indices"], edge
 * attrs "index"/"link") and the QRMSAEnv constructor arguments (envs/qrmsa.pyx:206-237).
 * Path p of the ordered pair (src,dst) has index (src*n_nodes + dst)*k_paths + p.
 */
typedef struct qrmsa_static_tables {
    int32_t n_nodes, n_links, k_paths, n_mods, mods_to_consider, n_rates, n_slots, max_hops;
    const uint8_t *path_hops;      /* [n_nodes*n_nodes*k_paths]            hops, 0 = no such path     */
    const uint8_t *path_links;     /* [n_nodes*n_nodes*k_paths*max_hops]   link indices               */
    const int32_t *link_n_spans;   /* [n_links]  spans per link (all spans of a link identical)        */
    const double *link_span_len_m; /* [n_links]  span length, metres                                   */
    const double *link_alpha;      /* [n_links]  Span.attenuation_normalized, 1/m (topology.pyx:21)    */
    const double *link_nf;         /* [n_links]  Span.noise_figure_normalized (topology.pyx:23)        */
    const int32_t *mod_se;         /* [n_mods]   Modulation.spectral_efficiency                        */
    const double *mod_min_osnr;    /* [n_mods]   Modulation.minimum_osnr, dB                           */
    const double *bit_rates;       /* [n_rates]  bit-rate classes (Gb/s)                               */
    const uint8_t *slots_needed;   /* [n_rates*n_mods] QRMSAEnv.get_number_slots (qrmsa.pyx:1198-1205) */
    double frequency_start;        /* Hz  (qrmsa.pyx:226)  */
    double slot_bandwidth_hz;      /* Hz  (qrmsa.pyx:227)  */
    double launch_power_w;         /* W   (qrmsa.pyx:288)  */
    double margin_db;              /* dB  (qrmsa.pyx:228)  */
    /* only read by qrmsa_observation (route-length feature, qrmsa.pyx:676-690); may be NULL otherwise */
    const double *path_length_km;  /* [n_nodes*n_nodes*k_paths] Path.length                            */
    const double *link_length_km;  /* [n_links] topology[u][v]["length"]                               */
} qrmsa_static_tables;

/* Counter slots returned by qrmsa_counters (per env group), all int64. */
enum {
    QRMSA_CNT_DECIDED = 0,        /* requests decided (= env.step calls that consumed a request)     */
    QRMSA_CNT_ACCEPTED = 1,       /* services_accepted            (qrmsa.pyx:1316-1317)              */
    QRMSA_CNT_REJECTED = 2,       /* bl_reject                    (qrmsa.pyx:861-865)                */
    QRMSA_CNT_RATE_REQUESTED = 3, /* sum of requested bit rates   (qrmsa.pyx:1106-1107)              */
    QRMSA_CNT_RATE_PROVISIONED = 4, /* sum of provisioned rates   (qrmsa.pyx:1318-1321)              */
    QRMSA_CNT_HOPS_ACCEPTED = 5,  /* sum of hops of accepted paths (h of SURVEY 8d)                  */
    QRMSA_CNT_LINKS_READ = 6,     /* link bitmap rows read while deciding (Lr)                       */
    QRMSA_CNT_RECORDS_READ = 7,   /* channel records on links of paths that had a QoT check (Nq)     */
    QRMSA_CNT_GN_TERMS = 8,       /* (link, neighbour) terms summed by the GN model                  */
    QRMSA_CNT_GN_EVALS = 9,       /* GSNR evaluations                                                */
    QRMSA_CNT_RELEASES = 10,      /* services released            (qrmsa.pyx:1113-1122)              */
    QRMSA_CNT_NEAR_THRESHOLD = 11, /* decisions with some |GSNR - threshold| < 1e-3 dB (flagged)     */
    QRMSA_CNT_BLOCKED_RESOURCES = 12, /* heuristic return flag    (heuristics.py:966)                */
    QRMSA_CNT_BLOCKED_OSNR = 13,
    QRMSA_CNT_PATHS_TRIED = 14,
    QRMSA_CNT_ERRORS = 15,        /* envs that hit an error state (list overflow, ValueError path)   */
    QRMSA_CNT_MOD_HIST = 16,      /* [16..23] accepted services per modulation index                 */
    QRMSA_CNT_GN_PRUNED = 24,     /* QoT checks refused on the empty-network bound without a GN sum   */
    QRMSA_CNT_DISRUPTED = 25,     /* disrupted_services           (qrmsa.pyx:937-952)                */
    QRMSA_CNT_DEFRAG_CYCLES = 26, /* episode_defrag_cicles        (qrmsa.pyx:1546)                   */
    QRMSA_CNT_REALLOCATIONS = 27, /* episode_service_realocations (qrmsa.pyx:1635)                   */
    QRMSA_N_COUNTERS = 32
};

/* Bits OR-ed into the per-request action word above bit 24 (see qrmsa_get_actions). */
#define QRMSA_ACTION_MASK 0x00ffffffu
#define QRMSA_FLAG_NEAR_THRESHOLD 0x80000000u /* some evaluated GSNR within 1e-3 dB of its threshold */
#define QRMSA_FLAG_DECIDED 0x40000000u        /* the request has been decided                        */
#define QRMSA_FLAG_ACCEPTED 0x20000000u
#define QRMSA_FLAG_BLOCKED_RESOURCES 0x01000000u /* rejected: the heuristic's blocked_due_to_resources (heuristics.py:966) */
#define QRMSA_FLAG_BLOCKED_OSNR 0x02000000u      /* rejected: blocked_due_to_osnr                                     */
#define QRMSA_FLAG_DISRUPTED 0x10000000u        /* in disrupted_services_list: GSNR fell below minimum_osnr (qrmsa.pyx:937-952) */
#define QRMSA_FLAG_RELEASE_CANCELLED 0x08000000u /* accepted, but its release event was dropped (qrmsa.pyx:433, :461-464)  */
#define QRMSA_FLAG_NEAR_TIE 0x04000000u          /* highest-SNR policy: a runner-up within 1e-6 dB of the chosen candidate */

/* step_action status per env (qrmsa.pyx:838-1065) */
enum {
    QRMSA_STEP_ACCEPTED = 0,
    QRMSA_STEP_REJECT_ACTION = 1, /* action == k*M*S                               (qrmsa.pyx:861-865) */
    QRMSA_STEP_NOT_FREE = 2,      /* is_path_free false: request NOT consumed      (qrmsa.pyx:886-897) */
    QRMSA_STEP_LOW_GSNR = 3,      /* reference raises ValueError                   (qrmsa.pyx:925-929) */
    QRMSA_STEP_IDLE = 4           /* env has no further request in the loaded trace                    */
};

const char *qrmsa_version(void);
const char *qrmsa_strerror(int status);
/* Last CUDA / validation message recorded on this context ("" if none). */
const char *qrmsa_last_error(const qrmsa_ctx *ctx);

/*
 * Replaces: QRMSAEnv.__init__ (envs/qrmsa.pyx:206-415) for n_envs independent environments.
 * Allocates all per-env state in device memory: packed per-link slot bitmaps (uint32 words, 1 =
 * free; stands in for topology.graph["available_slots"], qrmsa.pyx:302-305), per-link channel
 * lists (stand in for topology[u][v]["running_services"]), the request/service table and the
 * release schedule.  max_requests = requests per episode (episode_length), <= 16384.
 */
int qrmsa_create(const qrmsa_static_tables *tables, int n_envs, int max_requests, int device, qrmsa_ctx **out);
void qrmsa_destroy(qrmsa_ctx *ctx);

/* Env groups (contiguous, equal-sized) for per-load-point counters; default 1. */
int qrmsa_set_groups(qrmsa_ctx *ctx, int n_groups);
/*
 * What the step kernel keeps in shared memory beside the GN tables.  2 (default): per warp the env's link rows and the
 * request / schedule stream chunks plus a compact path table, whenever they fit (else as 1); 1: the stream chunks only
 * (else as 0); 0: nothing -- all env state through L1/L2.  The three variants compute the same words; lowering the level
 * is what a configuration with larger tables gets by itself, and is how tests hold the variants to each other.
 */
int qrmsa_set_staging(qrmsa_ctx *ctx, int level);
/*
 * Replaces: the constructor switches measure_disruptions / defragmentation / n_defrag_services (envs/qrmsa.pyx:206-237).
 *   measure_disruptions  after every accepted service, each running service on the links of its path that is not yet
 *                        marked disrupted is re-evaluated against all current channels; GSNR below its modulation's
 *                        minimum_osnr (no margin) marks it (QRMSA_FLAG_DISRUPTED) and counts it once (qrmsa.pyx:937-952)
 *   defragmentation      after every release -- every n_defrag_services-th processed request, or every one when 0 --
 *                        each running service, in provisioning order, moves to the lowest valid start below its own
 *                        where its GSNR still meets minimum_osnr (qrmsa.pyx:1113-1122, :1545-1639); the start slot in
 *                        its action word changes with it
 * Applies to qrmsa_step_action and to the first-fit policy of qrmsa_step_heuristic (a general-dimension kernel: these
 * switches cost 10-100x per request, as they do in the reference).  Not combinable with qrmsa_cancel_pending_releases.
 */
int qrmsa_set_features(qrmsa_ctx *ctx, int measure_disruptions, int defragmentation, int n_defrag_services);
/* Services found disrupted by the LAST decided request of every env (the disrupted_services column of the per-service
 * CSV, qrmsa.pyx:983): int32 [n_envs]; needs measure_disruptions. */
int qrmsa_get_step_disrupted_host(qrmsa_ctx *ctx, int32_t *h_out, void *stream);
/* Keep a per-request GSNR log (double, [n_envs][max_requests]); off by default. */
int qrmsa_enable_gsnr_log(qrmsa_ctx *ctx, int enable);

/*
 * Replaces: QRMSAEnv.reset (envs/qrmsa.pyx:427-504) for every env: all slots free, lists empty,
 * release queue cleared, episode counters zeroed.  The request stream of the new episode is
 * attached with qrmsa_load_trace*, whose request 0 becomes the current request (reset's
 * `_next_service()`, qrmsa.pyx:499-500).
 */
int qrmsa_reset(qrmsa_ctx *ctx, void *stream);

/*
 * Replaces: the `self._events = []` of QRMSAEnv.reset(options={"only_episode_counters": True}) (envs/qrmsa.pyx:433,
 * :461-464): the episode counters restart while the network is left as it is -- and, because the release heap is emptied,
 * every service that is running at that moment is never released.  Marks those services
 * (QRMSA_FLAG_RELEASE_CANCELLED in their action word); the request stream and the current request are untouched.
 */
int qrmsa_cancel_pending_releases(qrmsa_ctx *ctx, void *stream);

/*
 * Replaces: the per-request draws of QRMSAEnv._next_service/_get_node_pair
 * (envs/qrmsa.pyx:1079-1099, :1134-1148) by replaying a recorded stream, and the release heap of
 * _add_release (qrmsa.pyx:1327-1330) by a per-env release schedule sorted on the same key
 * (float32(arrival + holding), service_id).  Arrays are [n_requests][n_envs] (request-major);
 * n_requests <= max_requests.
 */
int qrmsa_load_trace(qrmsa_ctx *ctx, const uint8_t *d_src, const uint8_t *d_dst, const uint8_t *d_rate,
                     const float *d_arrival, const float *d_holding, int n_requests, void *stream);
int qrmsa_load_trace_host(qrmsa_ctx *ctx, const uint8_t *h_src, const uint8_t *h_dst, const uint8_t *h_rate,
                          const float *h_arrival, const float *h_holding, int n_requests, void *stream);

/*
 * Same as qrmsa_load_trace_host for a context that owns a SLICE of a larger batch: the host arrays are
 * [n_requests][row_stride] with row_stride >= n_envs, and the pointers address this context's first env.  Fully
 * asynchronous on `stream` (host memory must be pinned and stay alive); lets several contexts on several
 * streams overlap their uploads with each other's kernels.
 */
int qrmsa_load_trace_host_strided(qrmsa_ctx *ctx, const uint8_t *h_src, const uint8_t *h_dst, const uint8_t *h_rate,
                                  const float *h_arrival, const float *h_holding, int n_requests, int64_t row_stride,
                                  void *stream);

/*
 * Replaces: the same draws of QRMSAEnv._next_service / _get_node_pair (envs/qrmsa.pyx:1079-1099, :1134-1148) made ON
 * THE DEVICE for the next n_requests requests of every env: exponential inter-arrival and holding times
 * rounded to float32, the float32 clock, source / destination / bit rate by bisecting the cumulative-weight tables
 * (built by the caller exactly as for qrmsa_tracegen_create), drawn from a Philox4x32-10 stream keyed by
 * (seed, env_offset + env, request index).  The streams do not depend on how envs are split over contexts / GPUs;
 * they are not CPython's MT19937 streams (for replay parity with the reference use qrmsa_tracegen_* +
 * qrmsa_load_trace_host).  restart != 0 zeroes the clocks and the request counter; otherwise the streams continue
 * where the previous call stopped.  Also builds the release schedule, like qrmsa_load_trace.
 */
int qrmsa_generate_trace(qrmsa_ctx *ctx, uint64_t seed, int restart, int64_t env_offset, const double *h_load,
                         double mean_holding_time, const double *h_src_cum, const double *h_dst_cum,
                         const double *h_rate_cum, int n_requests, void *stream);
/* The loaded / generated request stream, requests [first, first+count), arrays [count][n_envs]; synchronises. */
int qrmsa_get_trace_host(qrmsa_ctx *ctx, int first, int count, uint8_t *h_src, uint8_t *h_dst, uint8_t *h_rate,
                         float *h_arrival, float *h_holding, void *stream);

/*
 * Replaces: n_steps iterations of the benchmark loop
 *     action, _, _ = heuristic_shortest_available_path_first_fit_best_modulation(env)
 *     env.step(action)
 * (heuristics/heuristics.py:923-966, envs/qrmsa.pyx:838-1065, examples/JOCN_Benchmark_2024/
 * graph_load.py:161-163) for every env, fused in one kernel launch: availability AND
 * (qrmsa.pyx:1482-1512), first-fit with guard band (utils.pyx:44-58, qrmsa.pyx:515-541), GN-model
 * GSNR (core/osnr.pyx:21-142), commit (qrmsa.pyx:1288-1330), next request + releases
 * (qrmsa.pyx:1067-1122, :1332-1350).  Steps past the end of the loaded trace are not executed.
 */
int qrmsa_step_first_fit(qrmsa_ctx *ctx, int n_steps, void *stream);

/*
 * Same fused loop with the heuristic chosen by id (the indices examples/JOCN_Benchmark_2024/graph_load.py:116-125
 * selects by number):
 *   QRMSA_POLICY_FIRST_FIT       heuristic_shortest_available_path_first_fit_best_modulation (heuristics.py:923-966)
 *   QRMSA_POLICY_LOAD_BALANCING  load_balancing_best_modulation (heuristics.py:547-627)
 *   QRMSA_POLICY_LB_FIRST_FIT    heuristic_load_balancing_first_fit (heuristics.py:202-270): paths ordered by the occupied
 *                                fraction of their availability, then first fit on the first path that admits a modulation
 *   QRMSA_POLICY_HIGHEST_SNR     heuristic_highest_snr (heuristics.py:272-328): every valid start of every (path,
 *                                modulation) is QoT-checked, the acceptable candidate with the highest GSNR wins
 */
enum { QRMSA_POLICY_FIRST_FIT = 0, QRMSA_POLICY_LOAD_BALANCING = 1, QRMSA_POLICY_HIGHEST_SNR = 2, QRMSA_POLICY_LB_FIRST_FIT = 3 };
int qrmsa_step_heuristic(qrmsa_ctx *ctx, int policy, int n_steps, void *stream);

/*
 * Replaces: QRMSAEnv.step(action) (envs/qrmsa.pyx:838-1065) with one externally chosen action per
 * env (RL path).  d_action int64[n_envs]; outputs (nullable) d_reward float[n_envs]
 * (qrmsa.pyx:992-995, :1266-1285), d_status uint8[n_envs] (QRMSA_STEP_*), d_gsnr double[n_envs],
 * d_terminated uint8[n_envs] (episode_services_processed == episode_length, qrmsa.pyx:1056).
 */
int qrmsa_step_action(qrmsa_ctx *ctx, const int64_t *d_action, float *d_reward, uint8_t *d_status, double *d_gsnr,
                      uint8_t *d_terminated, void *stream);

/*
 * Replaces: QRMSAEnv.observation() with gen_observation=True (envs/qrmsa.pyx:583-781, GSNR per candidate from
 * core.osnr.calculate_osnr_observation, core/osnr.pyx:259-368) for the CURRENT request of every env:
 *   d_obs   float [n_envs][1 + 2 + k + 12*k*M]   bit rate, source, destination, k route lengths, 12 features per
 *                                                (path, modulation) (qrmsa.pyx:592-665, :773-779)
 *   d_mask  uint8 [n_envs][k*M*S + 1]            1 iff the start slot is a guard-band-valid candidate AND
 *                                                round((gsnr - minimum_osnr)/|minimum_osnr|, 10) >= 0 (threshold
 *                                                WITHOUT margin); last entry (reject) always 1 (qrmsa.pyx:731-766)
 * Read-only on the env state.  Stands in for the mask SB3's MaskablePPO asks the wrapper for
 * (wrappers/qrmsa_gym.py:74-75).
 */
int qrmsa_observation(qrmsa_ctx *ctx, float *d_obs, uint8_t *d_mask, void *stream);
/*
 * modulations_to_consider < n_mods (examples/ONDM_2025/new_train_multi_ppo.py:101): the action space has
 * k * modulations_to_consider * S + 1 entries and block j of a path stands for modulation max_modulation_idx - j, where
 * max_modulation_idx is re-decided by every observation (QRMSAEnv.get_max_modulation_index, envs/qrmsa.pyx:543-581: the
 * best modulation with an acceptable candidate on the first path that has one, never below modulations_to_consider - 1)
 * and used by step() to decode the action (qrmsa.pyx:821-829).  uint8 [n_envs]: the value of every env's last observation
 * (n_mods - 1 after a reset).  The decision log keeps the absolute modulation: its action words are composed with n_mods.
 */
int qrmsa_get_max_modulation_idx_host(qrmsa_ctx *ctx, uint8_t *h_out, void *stream);
/* Sizes of one env's observation vector and action mask. */
int qrmsa_observation_dims(const qrmsa_ctx *ctx, int *obs_dim, int *n_actions);

/*
 * Replaces: the masked categorical sample sb3_contrib's MaskablePPO draws from the policy's logits and the wrapper's
 * action_masks() (wrappers/qrmsa_gym.py:74-75, examples/ONDM_2025/train_multi_masked_ppo.py:410-444), for every env in one
 * pass: d_action[e] ~ Categorical(softmax(d_logits[e]) restricted to d_mask[e] != 0), drawn by the Gumbel-max identity from
 * a Philox4x32-10 stream keyed by (seed, step) and counted by (env, action): reproducible, independent of the launch shape.
 * d_logits: [n_envs][logit_row_stride] float32 or bfloat16 (QRMSA_LOGITS_*); d_mask: uint8 [n_envs][mask_row_stride] as
 * written by qrmsa_observation; d_action: int64 [n_envs].  No context: the call only needs a device.
 */
enum { QRMSA_LOGITS_F32 = 0, QRMSA_LOGITS_BF16 = 1 };
int qrmsa_sample_masked_actions(const void *d_logits, int logits_dtype, const uint8_t *d_mask, int n_envs, int n_actions,
                                int64_t logit_row_stride, int64_t mask_row_stride, uint64_t seed, uint64_t step,
                                int64_t *d_action, int device, void *stream);

/*
 * Decisions of requests [first, first+count) as int32 action words, [count][n_envs] request-major:
 * low 24 bits = the reference's action index p*M*S + (max_mod_idx - m)*S + slot, reject = k*M*S
 * (heuristics.py:36-54, qrmsa.pyx:319-321), high bits = QRMSA_FLAG_*.
 */
int qrmsa_get_actions(qrmsa_ctx *ctx, int first, int count, int32_t *d_out, void *stream);
int qrmsa_get_actions_host(qrmsa_ctx *ctx, int first, int count, int32_t *h_out, void *stream);
/* Asynchronous, strided variant: h_out is [count][row_stride] int32 (pinned), this context's envs start at h_out. */
int qrmsa_get_actions_host_strided(qrmsa_ctx *ctx, int first, int count, int32_t *h_out, int64_t row_stride,
                                   void *stream);
/*
 * Request / service / decision records [first, first+count) of ONE env, uint32 [count][4] =
 * {arrival float32 bits, holding float32 bits, src | dst << 8 | rate index << 16, action word} -- the env's request
 * stream (Service fields of envs/qrmsa.pyx:29-116) and its decision log in one copy; used to replay sampled envs of a
 * large batch through a checker.  Synchronises the device.
 */
int qrmsa_get_env_log_host(qrmsa_ctx *ctx, int env, int first, int count, uint32_t *h_records4);
/* GSNR (dB) of the accepted candidate per request (0.0 on reject), needs qrmsa_enable_gsnr_log. */
int qrmsa_get_gsnr_host(qrmsa_ctx *ctx, int first, int count, double *h_out, void *stream);
/* ASE-only and NLI-only figures (dB) of the same candidates: the 2nd and 3rd value calculate_osnr returns
 * (core/osnr.pyx:133-142), written to the reference's per-service CSV (envs/qrmsa.pyx:967-990). */
int qrmsa_get_ase_nli_host(qrmsa_ctx *ctx, int first, int count, double *h_ase, double *h_nli, void *stream);

/* Per-group counters, int64 [n_groups][QRMSA_N_COUNTERS]; synchronises the stream. */
int qrmsa_counters(qrmsa_ctx *ctx, int64_t *h_out, void *stream);
/* Same, left on the device (for the episode-end NCCL all-reduce): int64 [n_groups][QRMSA_N_COUNTERS]. */
int qrmsa_counters_device(qrmsa_ctx *ctx, int64_t **d_out);
/* Per-env progress: int32 [n_envs][4] = {current request index, accepted, rejected, error code}. */
int qrmsa_env_state_host(qrmsa_ctx *ctx, int32_t *h_out, void *stream);

/*
 * Replaces: reading topology.graph["available_slots"] (int32 [E][S], 1 = free; qrmsa.pyx:302-305)
 * of one env -- parity / debugging / the single-env view used by the Python heuristics.
 */
int qrmsa_export_slots(qrmsa_ctx *ctx, int env, int32_t *h_available_slots);
/* Packed bitmaps of envs [first, first+count): uint32 [count][E][W], W = ceil(S/32). */
int qrmsa_export_bitmaps(qrmsa_ctx *ctx, int first, int count, uint32_t *h_out);
/* Channel list of one link (stand-in for topology[u][v]["running_services"]): int32 [cap][3] =
 * (initial_slot, number_slots, modulation index); returns the count in *n. */
int qrmsa_export_link_list(qrmsa_ctx *ctx, int env, int link, int32_t *h_out3, int cap, int *n);
/*
 * Replaces: core.osnr.calculate_osnr(env, service) (core/osnr.pyx:21-142) for a hypothetical
 * service on path p of (src,dst) at (initial_slot, number_slots) against env's current state.
 */
int qrmsa_probe_gsnr(qrmsa_ctx *ctx, int env, int src, int dst, int p, int initial_slot, int number_slots,
                     double *h_gsnr_db);
/* The same call with all three values calculate_osnr returns (core/osnr.pyx:133-142): h_out[0] = GSNR, h_out[1] = the
 * ASE-only figure, h_out[2] = the NLI-only figure, in dB. */
int qrmsa_probe_qot(qrmsa_ctx *ctx, int env, int src, int dst, int p, int initial_slot, int number_slots,
                    double *h_gsnr_ase_nli_db);

/*
 * Host-side request generator reproducing CPython 3.12 `random.Random(seed)` draw for draw
 * (MT19937, expovariate, choices) in the order of QRMSAEnv._next_service (qrmsa.pyx:1079-1089)
 * and _get_node_pair (qrmsa.pyx:1134-1148): env i is seeded with base_seed + i.
 * h_load[n_envs] = offered load per env; h_src_cum[n_nodes] / h_dst_cum[n_nodes*n_nodes] /
 * h_rate_cum[n_rates] are the cumulative weight tables `choices` bisects (built by the caller
 * exactly as the reference builds its weights).  Pure host code, no GPU needed.
 */
int qrmsa_tracegen_create(int n_envs, uint64_t base_seed, int n_nodes, int n_rates, const double *h_load,
                          double mean_holding_time, const double *h_src_cum, const double *h_dst_cum,
                          const double *h_rate_cum, qrmsa_tracegen **out);
/* bit_rate_selection="continuous" (qrmsa.pyx:246-254, :1086-1087): the bit rate is `rng.randint(lower, higher)` instead
 * of `choices(bit_rates, probs)`; the generated rate index is (bit rate - lower), i.e. an index into the table
 * (lower, lower+1, ..., higher), which must have at most 255 entries.  Call before the first qrmsa_tracegen_next. */
int qrmsa_tracegen_set_randint_rates(qrmsa_tracegen *gen, int lower, int higher);
/* Next n_requests of every env, arrays [n_requests][n_envs]; the clock persists across calls. */
int qrmsa_tracegen_next(qrmsa_tracegen *gen, int n_requests, uint8_t *h_src, uint8_t *h_dst, uint8_t *h_rate,
                        float *h_arrival, float *h_holding, int n_threads);
void qrmsa_tracegen_destroy(qrmsa_tracegen *gen);

#ifdef __cplusplus
}
#endif
#endif /* QRMSA_B200_H */
