/*
 * conclave_b200.h — C ABI of the B200-native conclave voting engine.
 *
 * The reference (potalora/conclave_sim) is pure Python and has no FFI of its
 * own; this header is the boundary a maintainer would bind (ctypes stub in
 * INTEGRATION.md) to replace, for the Monte Carlo hot path only:
 *
 *   csim_upload_roster   <- TransitionModel.__init__            src/model.py:172-254
 *                           (captures ideology_score / region / is_papabile, :199-213)
 *   csim_beta_schedule   <- TransitionModel.update_beta_weight  src/model.py:562-580
 *   csim_need_votes      <- the supermajority test              src/simulate.py:174-175
 *   csim_prob_matrix     <- TransitionModel.calculate_transition_probabilities
 *                                                               src/model.py:269-560
 *   csim_run             <- _run_single_simulation_worker       src/simulate.py:30-218
 *                           fanned out over trials like         src/simulate.py:330-351
 *                           + the histograms the aggregation    src/simulate.py:356-387
 *                           needs
 *   csim_group_run       <- the ProcessPoolExecutor fan-out     src/simulate.py:330-351
 *                           (`-w/--workers`, :231): the trials of one call sharded over the
 *                           GPUs of one box by disjoint sim-id ranges, histograms combined by
 *                           ONE NCCL all-reduce (csim_allreduce_hists)
 *   csim_engine_probs    <- the same model.py:269-560 matrix, but computed with the arithmetic
 *                           each TRIAL engine uses in its hot loop (parity probe: 1e-5 gate)
 *
 * Conventions: plain pointers and sizes, caller owns every buffer, no
 * exceptions cross the boundary.  Every call returns CSIM_OK (0) or a negative
 * csim_status; csim_last_error() gives the message of the calling thread's
 * last failure.  A context is bound to one CUDA device and owns staging
 * buffers that successive calls reuse: calls on ONE context must be issued
 * from one thread at a time; use one context per host thread (contexts are
 * independent).  A context serves ONE stream at a time: its per-batch scratch
 * (parameter sets, beta tables, CDF / prefix tables) is overwritten by the next
 * call, so every call first makes its stream wait (cudaStreamWaitEvent) on the
 * work the previous call of the same context enqueued — calls on different
 * streams are therefore serialised, never racing.  Every entry point restores
 * the calling thread's current CUDA device before it returns.  There is NO CPU
 * fallback: with no usable CUDA device csim_create fails with CSIM_ERR_CUDA.
 */
#ifndef CONCLAVE_B200_H
#define CONCLAVE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CSIM_ABI_VERSION 2

typedef enum csim_status {
    CSIM_OK = 0,
    CSIM_ERR_INVALID = -1,   /* bad argument */
    CSIM_ERR_CUDA = -2,      /* CUDA runtime error (message has the cudaError string) */
    CSIM_ERR_STATE = -3,     /* e.g. roster not uploaded */
    CSIM_ERR_UNSUPPORTED = -4
} csim_status;

/* wiring: which inputs of the model are connected */
#define CSIM_WIRING_REF_CLI 0 /* what `python -m src.simulate` computes: previous votes never reach the model */
#define CSIM_WIRING_MODEL   1 /* bandwagon / stickiness / fatigue / stop-candidate live, fed per trial */

/* engine */
#define CSIM_ENGINE_AUTO  0   /* ref_cli: TABLE when the batch is large enough to fill the GPU and E <= 256 (or fp64), else
                                 PREFIX; model: PREFIX, or FULL for what PREFIX refuses in fp32 (candidate fatigue / stop-candidate
                                 live, negative bonus / strength / papabile factors); fp64: TABLE (ref_cli) or DENSE (model).
                                 A shape the chosen engine cannot hold in shared memory falls through to FULL, then DENSE,
                                 then the fp64 exact engine (global-memory tables; results are then of CSIM_FP64 quality and
                                 ~100x slower): AUTO returns CSIM_ERR_UNSUPPORTED only for what no engine runs; an explicitly
                                 named engine reports its own limits.  csim_last_run_info says which engine / precision ran */
#define CSIM_ENGINE_DENSE 1   /* every trial re-evaluates its E x C scores each round */
#define CSIM_ENGINE_TABLE 2   /* ref_cli only: trial-invariant per-round CDF tables, built once */
#define CSIM_ENGINE_PREFIX 3  /* fp32, both wirings: trial-invariant chunk-prefix tables + per-trial bandwagon /
                                 stickiness corrections, binary search over chunks, re-scan of one chunk */
#define CSIM_ENGINE_FULL   4  /* fp32: every term of the model, candidate fatigue and stop-candidate included (the fatigue set
                                 and the stop trigger are decided in fp64 exactly as the exact engine decides them).  Two
                                 kernels behind it: with fatigue / stop-candidate live, E <= 255, non-negative factors and a
                                 factorisable beta schedule the EXT instantiation of the dense kernel (fatigue folded into a
                                 per-warp round table, stop-candidate as a byte-masked second accumulator); otherwise every
                                 term inline per (elector, candidate) with the model's clamps (negative factors, any E) */

/* arithmetic */
#define CSIM_FP32 0           /* fast path: fp32 scores, ex2.approx, 24-bit uniforms */
#define CSIM_FP64 1           /* exact path: fp64, numpy summation order, 53-bit uniforms;
                                 bit-exact against the reference under a shared uniform tape */

/* reason codes (strings of src/simulate.py:61,130,178) */
#define CSIM_REASON_MAX_ROUNDS 0 /* "Max rounds reached" */
#define CSIM_REASON_WINNER     1 /* "Winner found by supermajority (NN% threshold)" */
#define CSIM_REASON_NO_PROBS   2 /* "No probabilities generated" */

/* beta_mode: which branch of update_beta_weight is live (model.py:569-578) */
#define CSIM_BETA_STEPPED 0   /* beta_increment_amount is not None: beta += amount when r % interval == 0 */
#define CSIM_BETA_GROWTH  1   /* beta_increment_amount is None and growth_rate > 1: beta *= rate */

/* One parameter set = TransitionModel kwargs (model.py:172-191) + worker arguments
 * (simulate.py:30-41).  fp64 like the reference; all fields plain scalars. */
typedef struct csim_params {
    double  initial_beta_weight;
    double  beta_increment_amount;
    double  beta_growth_rate;
    double  stickiness_factor;
    double  bandwagon_strength;
    double  regional_affinity_bonus;
    double  papabile_weight_factor;
    double  fatigue_vote_share_threshold;
    double  fatigue_penalty_factor;
    double  stop_candidate_threshold_unacceptable_distance;
    double  stop_candidate_threat_min_vote_share;
    double  stop_candidate_boost_factor;
    double  supermajority_threshold;
    int32_t enable_dynamic_beta;
    int32_t beta_mode;
    int32_t beta_increment_interval_rounds;
    int32_t enable_candidate_fatigue;
    int32_t fatigue_threshold_rounds;
    int32_t fatigue_top_n_immune;
    int32_t enable_stop_candidate;
    int32_t max_rounds;
    int32_t runoff_threshold_rounds;
    int32_t runoff_candidate_count;
    int32_t reserved0;
    int32_t reserved1;
} csim_params;

typedef struct csim_ctx csim_ctx;

/* Arguments of one batch of trials.  Trials are laid out [set][sim]:
 * trial t = set * n_sims + i simulates parameter set `set` with
 * sim_id = sim_id_begin + i.  Under the native RNG (uniform_tape == NULL) the
 * votes of a trial depend only on (seed, sim_id, params): results do not
 * depend on batching, sharding or device.
 * Limit of a batch: every parameter set must have the same max_rounds (the per-trial
 * outputs and rounds_hist are dense [.., max_rounds, ..] arrays); a sweep over
 * max_rounds is one csim_run per distinct value (CSIM_ERR_INVALID otherwise).
 * n_param_sets * n_sims == 0 is a valid empty batch: the call returns CSIM_OK
 * without touching any buffer (pointers may then be NULL). */
typedef struct csim_run_args {
    const csim_params* params;      /* [n_param_sets] (host memory, always) */
    int32_t  n_param_sets;
    int32_t  wiring;                /* CSIM_WIRING_* */
    int32_t  engine;                /* CSIM_ENGINE_* */
    int32_t  precision;             /* CSIM_FP32 / CSIM_FP64 */
    uint64_t sim_id_begin;
    uint64_t n_sims;                /* per parameter set */
    uint64_t seed;                  /* Philox4x32-10 key; counter = (sim_lo, sim_hi, round, elector>>2 [| 1<<31]), word elector&3 */
    const double* uniform_tape;     /* nullable: [n_trials, max_rounds * E], E uniforms per round in
                                       elector order — "the identical uniform draws exported from the
                                       reference".  Requires precision == CSIM_FP64. */
    /* per-trial outputs; n_trials = n_param_sets * n_sims */
    int32_t* winner;                /* [n_trials]  candidate index, -1 = none              (required) */
    int32_t* rounds;                /* [n_trials]  round of the win, or max_rounds          (required) */
    uint8_t* reason;                /* [n_trials]  CSIM_REASON_*                            (nullable) */
    int32_t* final_votes;           /* [n_trials, E]  last round's tally                    (nullable) */
    int32_t* round_tallies;         /* [n_trials, max_rounds, E], -1 beyond `rounds`        (nullable) */
    /* aggregated outputs, ACCUMULATED into (caller zeroes them) */
    int64_t* winner_hist;           /* [n_param_sets, E + 1]; slot E = no winner            (nullable) */
    int64_t* rounds_hist;           /* [n_param_sets, max_rounds + 1] over winning trials   (nullable) */
    int64_t* totals;                /* [n_param_sets, 2]: {sum of rounds played, trials}    (nullable) */
    int32_t  buffers_on_device;     /* 0: every pointer above is host memory; the call copies, runs and
                                          synchronises.  1: device pointers of ctx's device; the call only
                                          enqueues work on `stream` (uniform_tape included). */
    int32_t  reserved;
    void*    stream;                /* cudaStream_t, NULL = default stream */
    uint64_t auto_batch_sims;       /* 0, or the trials per parameter set of the WHOLE job this batch is one shard of:
                                       CSIM_ENGINE_AUTO then chooses as it would for the whole job, so that every shard of a
                                       multi-GPU run — and the single-device run it must reproduce — use the same engine
                                       (AUTO's choice depends on the batch size).  csim_group_run sets it for its devices. */
} csim_run_args;

int         csim_abi_version(void);
const char* csim_last_error(void);

/* number of CUDA devices visible to the process (0 and CSIM_ERR_CUDA when the driver reports none) */
csim_status csim_device_count(int32_t* n);
csim_status csim_create(csim_ctx** out, int device);
csim_status csim_destroy(csim_ctx* ctx);
csim_status csim_device_info(csim_ctx* ctx, int* sm_count, int* cc_major, int* cc_minor, int* clock_khz);

/* Roster: uploaded once.  region = arbitrary integer codes (equal code == same region). */
csim_status csim_upload_roster(csim_ctx* ctx, int32_t n_electors, const double* ideology_score,
                               const int32_t* region_code, const uint8_t* is_papabile);

/* Host-side, fp64, bit-identical to the reference's repeated add / multiply:
 * out[i] = beta in effect in round first_round + i when the model is called once per round. */
csim_status csim_beta_schedule(const csim_params* p, int32_t first_round, int32_t n_rounds, double* out);
/* smallest integer v with (double)v >= threshold * n_electors */
int32_t     csim_need_votes(double supermajority_threshold, int32_t n_electors);
/* smallest vote count v in [0, n_electors + 1] with (double)v / n_electors >= share (strict == 0) or > share (strict != 0):
 * the share tests of candidate fatigue (simulate.py:111: share < threshold) and stop-candidate (model.py:397: share >
 * threshold) as exact integer tests on vote counts */
int32_t     csim_vote_threshold(double share, int32_t n_electors, int32_t strict);

/* P[E, n_cand] for one round with effective beta `beta` (host buffers in and out).
 * prev_vote[E]  : candidate index (roster order) each elector voted for last round, -1 = none (nullable -> stickiness off)
 * prev_count[E] : last round's count per candidate                              (nullable -> bandwagon off)
 * fatigued[E]   : 1 = candidate fatigued (used iff p->enable_candidate_fatigue) (nullable)
 * shares[E]     : last round's vote shares (used iff p->enable_stop_candidate)  (nullable)
 * cand_index[n_cand] : the "effective candidates" of model.py:280-296 as roster indices, in output column
 *                 order; NULL -> all E candidates (n_cand ignored).  Rows are normalised over these columns only;
 *                 the caller zeroes prev_count / prev_vote entries of candidates outside the subset
 *                 (model.py:337,359 filter them the same way).
 * precision     : CSIM_FP64 -> P_out is double*, CSIM_FP32 -> float*. */
csim_status csim_prob_matrix(csim_ctx* ctx, const csim_params* p, double beta,
                             const int32_t* prev_vote, const int32_t* prev_count,
                             const uint8_t* fatigued, const double* shares,
                             const int32_t* cand_index, int32_t n_cand,
                             int32_t precision, void* P_out);

/* The hot path: a batch of independent trials. */
csim_status csim_run(csim_ctx* ctx, const csim_run_args* args);

/* How the PREFIX engine would lay one launch out (pure host arithmetic, no device needed): used by the CPU tests to
 * check the shared-memory budgeting for any roster shape, and by callers that want to know whether a roster's tables
 * are staged in shared memory or probed through L2.
 *   table_rounds = rounds that can be full-field: min(max_rounds, max(runoff_threshold_rounds, 1)) when a run-off
 *                  exists, else max_rounds;  smem_budget = opt-in shared memory per block (232448 on sm_100).
 * offsets[] = {o_xs, o_pw, o_rb, o_xg, o_nbs (from the start of dynamic shared memory), w_bwp, w_tally, w_eff,
 *              w_votes, w_mark (from the warp's block), candidate-array pitch in bytes, 4 x first search step}. */
typedef struct csim_prefix_plan {
    int32_t  staged;            /* 1: tables of all table_rounds live in shared memory */
    int32_t  chunk;             /* candidates per chunk (8, 16, 32, ...) */
    int32_t  n_chunks;          /* <= 32 */
    int32_t  row_pitch;         /* floats per table row (odd) */
    int32_t  rows;              /* table rows per round */
    int32_t  table_rounds;
    int32_t  warps_per_cta;
    int32_t  e_quad_end;        /* electors below this are dealt as whole Philox quads per lane */
    uint64_t table_bytes_per_set;
    uint64_t smem_tables, smem_cta, smem_per_warp, smem_total;
    uint32_t offsets[12];
} csim_prefix_plan;
csim_status csim_prefix_plan_query(int32_t n_electors, int32_t n_regions, int32_t max_rounds, int32_t table_rounds, int32_t kmax,
                                   int32_t wiring, uint64_t smem_budget, int32_t allow_stage, csim_prefix_plan* out);
/* table row of an elector inside one round's table (the order the lanes visit electors) */
int32_t     csim_prefix_slot(int32_t elector, int32_t e_quad_end);

/* Which of its two kernels the FULL engine would run for a batch with candidate fatigue / stop-candidate live, and its
 * shared-memory budget (pure host arithmetic, no device needed; used by the CPU tests for any roster shape).
 *   full_field_rounds   = rounds that can be full-field: min(max_rounds, max(runoff_threshold_rounds, 1)) when a run-off
 *                         exists, else max_rounds;  fatigue_window = largest fatigue_threshold_rounds of the batch
 *   fast_path_eligible  = 1 when every factor is non-negative and the beta schedule is factorisable (beta >= 0 and
 *                         beta x half-range of the ideology scores <= 26): the caller's facts, not the roster's shape
 * The fast path (the EXT instantiation of the dense kernel) is chosen when eligible, 2 <= electors <= 255 and at least 10
 * of its warps fit per SM (2 when the general kernel k_trials_full32 does not fit at all). */
typedef struct csim_full_plan {
    int32_t  fast_path;         /* 1: the dense kernel's EXT instantiation; 0: k_trials_full32 */
    int32_t  general_fits;      /* k_trials_full32 fits in shared memory */
    int32_t  warps_per_cta;     /* of the kernel that would run (one CTA per SM) */
    int32_t  cta_tabs;          /* fast path: the round tables of a parameter set are CTA-shared (1) or staged per warp per round (0) */
    int32_t  n_chunks;          /* chunks of candidates per row */
    int32_t  acc_pitch;         /* fast path: entries per chunk row of the acceptable-bytes table */
    uint64_t smem_cta, smem_per_warp, smem_total;
} csim_full_plan;
csim_status csim_full_plan_query(int32_t n_electors, int32_t n_regions, int32_t max_rounds, int32_t full_field_rounds, int32_t kmax,
                                 int32_t fatigue_window, int32_t fast_path_eligible, csim_full_plan* out);

/* Page-locked host memory (cudaHostAlloc / cudaFreeHost) for the per-trial result arrays of a host-buffer csim_run:
 * copies into it run at full PCIe speed.  Optional: any host memory works. */
csim_status csim_host_alloc(uint64_t bytes, void** out);
csim_status csim_host_free(void* p);

/* Which engine the last successful csim_run of this context launched (CSIM_ENGINE_DENSE ... CSIM_ENGINE_FULL, never AUTO),
 * the precision it computed in (CSIM_FP64 when AUTO fell through to the exact engine although CSIM_FP32 was asked for), and
 * whether AUTO's first choice was refused (fell_through != 0).  Any pointer may be NULL. */
csim_status csim_last_run_info(csim_ctx* ctx, int32_t* engine, int32_t* precision, int32_t* fell_through);

/* ---------------------------------------------------------------------------------------------
 * Parity probe: the probabilities a TRIAL engine uses in its hot loop (not the standalone csim_prob_matrix kernels).
 * One full-field round of one parameter set, computed by a probe kernel that calls the engine's own __device__ routines
 * (dense: chunk_acc / probe32 / the re-scan of finish_draw on the k_round_tables32 vectors; prefix: k_prefix_tables32 rows +
 * the re-scan of prefix_draw; full: full32_score8; table: the fp32 CDF rows k_trials_table searches).
 *   engine      CSIM_ENGINE_DENSE / TABLE / PREFIX / FULL (not AUTO); precision must be CSIM_FP32
 *   round       1-based: the beta in effect is csim_beta_schedule(p, round, 1)
 *   prev_vote / prev_count   as csim_prob_matrix (model wiring; NULL = first round).  TABLE: must be NULL (ref_cli only).
 *                            Stop-candidate (FULL) takes last round's shares as prev_count / E, exactly as a trial does
 *   fatigued                 FULL only: the candidate-fatigue mask to apply (used iff p->enable_candidate_fatigue), else NULL
 *   scores_out  [E, E]   float: the per-(elector, candidate) score of the engine's re-scan (TABLE: cdf[e][c] - cdf[e][c-1])
 *   prefix_out  [E, n_chunks] float: the chunk prefixes the engine's binary search probes, in score units; the last one is
 *               the row sum S_e the target u*S_e is formed from (TABLE: n_chunks == E, the CDF row itself, S_e == last == 1)
 *   n_chunks_out, chunk_out  chunks per row and candidates per chunk of that layout
 * P[e][c] = scores_out[e][c] / prefix_out[e][n_chunks - 1]. */
csim_status csim_engine_probs(csim_ctx* ctx, const csim_params* p, int32_t wiring, int32_t engine, int32_t round,
                              const int32_t* prev_vote, const int32_t* prev_count, const uint8_t* fatigued,
                              float* scores_out, float* prefix_out, int32_t prefix_capacity, int32_t* n_chunks_out, int32_t* chunk_out);

/* ---------------------------------------------------------------------------------------------
 * Multi-GPU (replaces the process pool of src/simulate.py:330-351).  Trials shard by disjoint sim-id ranges;
 * the only exchange is ONE all-reduce (sum, int64) of the winner / rounds histograms and totals.
 * NCCL is bound at RUN time (dlopen), so the library loads and every single-GPU call works on a box without NCCL. */

/* dlopen NCCL: `path` (e.g. the wheel-bundled .../nvidia/nccl/lib/libnccl.so.2) or NULL -> "libnccl.so.2" by the loader's
 * search path ($CSIM_NCCL_LIB first).  Idempotent; the multi-GPU calls below load it on first use. */
csim_status csim_nccl_load(const char* path);
/* NCCL_VERSION_CODE of the bound library, 0 when not loaded */
int32_t     csim_nccl_version(void);

/* One process per GPU (torchrun / MPI style): rank 0 makes an id, the host program broadcasts its 128 bytes by whatever
 * means it has, every rank creates its communicator on its context's device.  `comm` is an ncclComm_t. */
#define CSIM_NCCL_ID_BYTES 128
csim_status csim_nccl_unique_id(uint8_t id[CSIM_NCCL_ID_BYTES]);
csim_status csim_nccl_comm_create(csim_ctx* ctx, const uint8_t id[CSIM_NCCL_ID_BYTES], int32_t n_ranks, int32_t rank, void** comm);
csim_status csim_nccl_comm_destroy(void* comm);
/* In-place sum over the ranks of `comm` (any ncclComm_t of ctx's device — one made above or the host program's own) of this
 * rank's histograms (DEVICE buffers, the ones a device-buffer csim_run accumulated into), enqueued on `stream` after the
 * trial kernel.  ONE ncclAllReduce when the three buffers are contiguous in the order winner_hist | rounds_hist | totals
 * (pass the first pointer and the other two where they lie), otherwise one grouped call.  rounds_hist / totals may be NULL. */
csim_status csim_allreduce_hists(csim_ctx* ctx, void* comm, int64_t* winner_hist, int64_t* rounds_hist, int64_t* totals,
                                 int32_t n_param_sets, int32_t max_rounds, void* stream);

/* One process, several GPUs: a group owns one context (and one stream, one NCCL communicator) per device. */
typedef struct csim_group csim_group;
csim_status csim_group_create(csim_group** out, const int32_t* devices, int32_t n_devices);
csim_status csim_group_destroy(csim_group* g);
int32_t     csim_group_size(const csim_group* g);
csim_ctx*   csim_group_ctx(csim_group* g, int32_t i);                       /* the i-th device's context (borrowed) */
csim_status csim_group_upload_roster(csim_group* g, int32_t n_electors, const double* ideology_score,
                                     const int32_t* region_code, const uint8_t* is_papabile);
/* [begin, begin + count) of shard `rank` of `n` trials split over `n_ranks` contiguous, balanced ranges */
void        csim_shard_range(uint64_t n, int32_t rank, int32_t n_ranks, uint64_t* begin, uint64_t* count);
/* csim_run over every device of the group.  HOST buffers only (args->buffers_on_device must be 0; args->stream is ignored).
 * Device i simulates sims [begin_i, begin_i + count_i) of args->n_sims (csim_shard_range) of EVERY parameter set, all devices
 * concurrently; the per-device histograms are summed by one NCCL all-reduce enqueued behind each device's trial kernel; the
 * per-trial rows are copied straight into the caller's [set][sim] arrays (page-locked arrays make those copies asynchronous).
 * Results are bit-identical to csim_run on one device with the same arguments (a trial's stream is keyed by its sim id).
 * kernel_ms (nullable, [n_devices]): device time of each device's trial kernel. */
csim_status csim_group_run(csim_group* g, const csim_run_args* args, float* kernel_ms);

/* Diagnostics for bench.py: number of kernel launches this context has issued,
 * and the device time (ms, CUDA events on the launch stream) of the last csim_run's trial kernel. */
csim_status csim_stats(csim_ctx* ctx, int64_t* kernel_launches, float* last_trial_kernel_ms);

/* Instruction-throughput microbenchmarks used as roofline denominators:
 * kind 0 = FP32 FFMA, 1 = MUFU.EX2, 2 = the MUFU pair body (FADD, FMUL, MUFU.EX2, FFMA),
 * 3 = the factorised packed pair body (2 FMUL2 + 2 FMNMX + FADD2 per two pairs; 8 pairs = 20 instr).
 * Returns giga thread-instructions per second over the whole device. */
csim_status csim_microbench(csim_ctx* ctx, int32_t kind, double* ginstr_per_s);

#ifdef __cplusplus
}
#endif
#endif /* CONCLAVE_B200_H */
