/*
 * c4a0_engine.h — C-ABI of the B200-native self-play engine (libc4a0_engine.so).
 *
 * This is the drop-in boundary for the reference's batched MCTS self-play path.  In the reference
 * the Python-facing function `c4a0_rust.play_games` (rust/src/pybridge.rs:20-53) calls
 * `self_play::self_play` (rust/src/self_play.rs:39-129), which spawns one NN thread and ncpu-1 MCTS
 * threads that push per-game pointer trees (rust/src/mcts.rs:27-32, 332-355) through channels.  Here
 * the same work is a handle to a set of games resident in HBM that advance in lockstep: one call
 * to c4a0_engine_step() performs, for every live game, what one trip through
 * `MctsThread::loop_once` (self_play.rs:268-323) + `MctsGame::on_received_policy` (mcts.rs:83-108)
 * does for one game.  The neural network stays outside (PyTorch): the engine writes the leaf
 * positions as NN input planes (c4r.rs:378-392, pybridge.rs:202-221) into a caller-owned device
 * buffer and reads the network's logits / q values from caller-owned device buffers, so both sides
 * share memory with no copies.
 *
 * Conventions: plain C types only; every function returns 0 on success or a negative C4A0_E_* code
 * and records a message retrievable with c4a0_last_error() (thread-local).  Nothing aborts the
 * process (the reference panics: pybridge.rs:30, 182-188).  One host thread per handle; one handle
 * per GPU (several handles per GPU are allowed and are how two half-batches ping-pong).  Pointers
 * named *_dev are device pointers on the engine's GPU; all others are host pointers.  `stream` is a
 * cudaStream_t passed as void* (0 = legacy default stream); step()/eval_builtin() only enqueue
 * kernels and are safe to capture into a CUDA graph.
 */
#ifndef C4A0_ENGINE_H
#define C4A0_ENGINE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define C4A0_ABI_VERSION 3
#define C4A0_N_ROWS 6               /* lib.rs:29 */
#define C4A0_N_COLS 7               /* lib.rs:28 */
#define C4A0_BUF_N_CHANNELS 2       /* lib.rs:30 */
#define C4A0_PLANE_LEN 84           /* c4r.rs:50-52: 2*6*7 */
#define C4A0_MAX_SAMPLES 43         /* 42 moves + the terminal position (mcts.rs:271-313) */

enum {
  C4A0_OK = 0,
  C4A0_E_INVALID = -1,  /* bad argument / call order */
  C4A0_E_CUDA = -2,     /* a CUDA runtime call failed; see c4a0_last_error() */
  C4A0_E_NOMEM = -3,
  C4A0_E_ENGINE = -4    /* device-side invariant violated (reference would panic) */
};

enum { C4A0_PLANES_F32 = 0, C4A0_PLANES_BF16 = 1 };

/* Synthetic evaluators that stand in for the network in parity tests and kernel benchmarks
 * (the reference's tests do the same: mcts.rs:469-485, self_play.rs:386-403). */
/* C4A0_EVAL_HASH_FLAT: the hash evaluator scaled to what a freshly initialised network answers (logits
 * within +-0.25, values within +-0.1): broad trees, for profiling the tree kernels at the bench's shape. */
enum { C4A0_EVAL_UNIFORM = 0, C4A0_EVAL_HASH = 1, C4A0_EVAL_HASH_FLAT = 2 };

/* Row (= slot) states as reported by c4a0_engine_fetch_rows(). */
enum { C4A0_ROW_IDLE = 0, C4A0_ROW_WAIT_NN = 1, C4A0_ROW_CONTINUE = 2, C4A0_ROW_NEED_MOVE = 3 };

typedef struct c4a0_engine c4a0_engine;

/* Arguments of self_play() (self_play.rs:39-46) plus the lockstep width. */
typedef struct {
  uint32_t n_slots;            /* games resident at once == rows of the NN batch (<= max_nn_batch_size) */
  uint32_t max_requests;       /* capacity of set_requests() */
  uint32_t n_mcts_iterations;  /* >= 1 */
  float c_exploration;
  float c_ply_penalty;
  uint32_t plane_dtype;        /* C4A0_PLANES_F32 | C4A0_PLANES_BF16 */
  uint32_t max_inline_sims;    /* simulations that need no network row (terminal leaf, evaluation-cache hit) a game may
                                  run inside one step (0 = default: 2, or 3 with the evaluation cache) */
  int32_t device;              /* CUDA device ordinal */
  uint32_t plane_stride;       /* elements between consecutive rows of planes_dev (0 = 84; a multiple of
                                  4, >= 84; elements 88.. of a row are never written; bf16 rows with a
                                  16-byte-aligned stride get zeros in elements 84..87) */
  uint32_t flags;              /* C4A0_FLAG_* */
  uint32_t arena_blocks;       /* tree blocks (160 B) per arena half per game; 0 = n_mcts_iterations + 2,
                                  the minimum.  Larger halves make re-rooting copy-free until a half fills. */
  uint32_t eval_cache_entries; /* with C4A0_FLAG_EVAL_CACHE: entries (64 B each) of the evaluation cache, rounded
                                  up to a power of two; 0 = sized from n_slots * n_mcts_iterations and the
                                  free device memory */
  uint32_t spec_rows;          /* with C4A0_FLAG_SPECULATE: rows a small batch is topped up to (0 = 8192, at most n_slots) */
  /* Dirichlet noise on the root's priors (AlphaZero's exploration noise; the reference has none — mcts.rs:114-132
   * stores the network's masked softmax unchanged — so both default to 0 = off, which is the parity path).
   * With alpha > 0 and epsilon > 0 a node's priors become (1 - epsilon) P + epsilon eta, eta ~ Dir(alpha) over its
   * legal moves, once, when the node becomes the root of a search (game start, every move): at its expansion if it
   * is expanded as the root, at the re-root (mcts.rs:187-206) if the subtree is reused.  The draw is a pure function
   * of (game_id, moves played, column), so games are reproducible and independent of scheduling. */
  float dirichlet_alpha;
  float dirichlet_epsilon;
} c4a0_config;

/* Evaluate every waiting leaf even when several games wait on the same (position, model); by default
 * equal leaves share one network row, as the reference's NN thread does (self_play.rs:203-208). */
#define C4A0_FLAG_NO_DEDUP 1u
/* Keep every network answer of the current set_requests() job in a device table keyed by (position,
 * model) and answer later requests for the same key from it, inside the tick, instead of sending the
 * leaf to the network again (transpositions within a tree, the same position reached by other games
 * or again after a move).  Off by default because it assumes what the reference does not: that the
 * evaluator is a pure function of (model, position) for the duration of the job.  With such an
 * evaluator the games are identical with and without the cache; only n_rows per tick and the
 * number of ticks shrink.  The table is emptied by every set_requests(). */
#define C4A0_FLAG_EVAL_CACHE 2u
/* Needs C4A0_FLAG_EVAL_CACHE.  While a tick's batch is small (the first plies, when all games share a
 * few positions, and the tail of a job, when few games are left) the network runs far below capacity
 * and a tick costs the same whatever its rows.  With this flag batches for which the games ask for
 * at most min(spec_rows / 2, 1024) rows are topped up to spec_rows with the children of the leaves that are being expanded; the answers go into the
 * evaluation cache, so that selection finds them answered when it gets there, and a game may then run
 * several simulations per tick (4 x max_inline_sims).  Same assumption and same guarantee as the
 * cache: game records do not change.  Rows of these evaluations have row_slot 0xffffffff.  The I/O
 * buffers must hold c4a0_engine_io_rows() = n_slots + spec_rows rows. */
#define C4A0_FLAG_SPECULATE 4u

typedef struct {
  uint32_t n_requests;   /* games submitted */
  uint32_t n_started;    /* games that have been given a slot */
  uint32_t n_finished;   /* games whose samples are complete */
  uint32_t n_running;    /* slots holding a live game */
  uint32_t n_movers;     /* games that moved in the last step */
  uint32_t n_rows;       /* network rows the last step packed: rows [0, n_rows) of planes_dev are live */
  int32_t error;         /* 0, or C4A0_E_ENGINE if a game hit a state where the reference panics */
} c4a0_progress;

/* Counters in the reference's own units (self_play.rs:352-381 shows the same three live). */
typedef struct {
  uint64_t sims;               /* on_received_policy equivalents actually executed */
  uint64_t nn_evals;           /* rows handed to the network (unique leaf positions per tick) */
  uint64_t leaf_requests;      /* games that waited for a network answer, summed over ticks */
  uint64_t terminal_leaf_sims; /* sims whose leaf was terminal (run inside the kernel, no NN row) */
  uint64_t skipped_root_sims;  /* sims the reference would still spend on a terminal root */
  uint64_t moves;
  uint64_t samples;
  uint64_t select_depth_sum;   /* sum over sims of the depth of the selected leaf */
  uint64_t expansions;         /* nodes expanded (== non-terminal leaves applied) */
  uint64_t steps;              /* c4a0_engine_step() calls since set_requests() */
  uint64_t compacted_blocks;   /* tree blocks copied when an arena half filled up */
  uint64_t compactions;        /* number of such copies (re-roots that fit the arena copy nothing) */
  uint64_t cache_hits;         /* leaves answered from the evaluation cache (C4A0_FLAG_EVAL_CACHE) */
  uint64_t cache_inserts;      /* network answers stored in it */
  uint64_t spec_rows;          /* rows evaluated ahead of time (C4A0_FLAG_SPECULATE); part of nn_evals */
} c4a0_stats;

const char *c4a0_last_error(void);
int c4a0_abi_version(void);

/* ---- engine lifetime ------------------------------------------------------------------- */
int c4a0_engine_create(const c4a0_config *cfg, c4a0_engine **out);
void c4a0_engine_destroy(c4a0_engine *e);
/* bytes of device memory the engine holds (arenas + state + sample store) */
size_t c4a0_engine_device_bytes(const c4a0_engine *e);
/* rows the NN I/O buffers must have: n_slots, or n_slots + spec_rows with C4A0_FLAG_SPECULATE (every
 * live game can ask for a row of its own on top of the speculative ones) */
uint32_t c4a0_engine_io_rows(const c4a0_engine *e);

/* NN I/O buffers, caller-owned device memory (DLPack / torch tensors):
 *   planes_dev : [io_rows][plane_stride] f32 or bf16 (cfg.plane_dtype); the first 84 elements of a
 *                row are the [2][6][7] planes                          <- pybridge.rs:202-221
 *   logits_dev : [io_rows][7] f32, q_penalty_dev / q_no_penalty_dev : [io_rows] f32   (io_rows =
 *                c4a0_engine_io_rows(): n_slots unless speculation is on)
 *                                                                      <- pybridge.rs:175-196
 * Rows are dense: after set_requests()/step() rows [0, n_rows) hold the distinct leaf positions that
 * wait for an answer (n_rows <= io_rows, reported by poll()), in no particular order; the network
 * must fill the same rows of the three output buffers before the next step().  Rows >= n_rows are
 * ignored. */
int c4a0_engine_bind_io(c4a0_engine *e, void *planes_dev, const float *logits_dev,
                        const float *q_penalty_dev, const float *q_no_penalty_dev);

/* The games to play: GameMetadata{game_id, player0_id, player1_id} (types.rs:36-48) as three host
 * arrays.  Resets all state; games [0, min(n, n_slots)) are seated and their root planes written
 * (self_play.rs:55-58).  Requires bind_io(). */
int c4a0_engine_set_requests(c4a0_engine *e, const uint64_t *game_id, const uint64_t *player0_id,
                             const uint64_t *player1_id, uint32_t n, void *stream);

/* One lockstep tick: every game consumes its network answer (mask+softmax, expand, negamax
 * backup: mcts.rs:83-155), plays its move if the root reached n_mcts_iterations (temperature +
 * seeded sample + re-root: self_play.rs:283-300, mcts.rs:187-222), is emitted if that ended it
 * (mcts.rs:271-313; the next waiting request takes its slot) and selects its next leaf
 * (mcts.rs:160-183); the distinct waiting leaves are then packed as rows [0, n_rows) of planes_dev. */
int c4a0_engine_step(c4a0_engine *e, void *stream);

/* step() with CUDA events around its kernels (synchronises): device milliseconds of k_step (the
 * whole tick of every game) and of k_tail (compaction bursts; ~0 otherwise).  Not capturable. */
int c4a0_engine_step_timed(c4a0_engine *e, void *stream, float *ms_step_kernel, float *ms_tail_kernel);

/* Developer aid: runs one step() with cycle counters inside the tick kernel and returns 8 words per
 * slot {load cycles, apply cycles, simulations, run cycles (moves + selection + terminal backups),
 * terminal simulations, leaf depth, store cycles, total cycles} (zeros for idle slots).  The cycle
 * figures are those of the game's warp.  Synchronises. */
int c4a0_engine_debug_phases(c4a0_engine *e, void *stream, uint32_t *out8_per_slot);

/* Fill logits/q buffers for the current leaves with a synthetic evaluator (parity tiers E0/E1). */
int c4a0_engine_eval_builtin(c4a0_engine *e, int kind, void *stream);

/* Synchronises `stream` and reports progress. */
int c4a0_engine_poll(c4a0_engine *e, c4a0_progress *out, void *stream);
int c4a0_engine_stats(c4a0_engine *e, c4a0_stats *out, void *stream);

/* Per-row view for the numpy-callback compatibility path and for tests: *n_rows, and for each live
 * row the leaf position and the model that has to evaluate it (mcts.rs:70-76).  The arrays must hold
 * c4a0_engine_io_rows() entries; any of them may be NULL. */
int c4a0_engine_fetch_rows(c4a0_engine *e, uint32_t *n_rows, uint64_t *leaf_mask,
                           uint64_t *leaf_value, uint64_t *model_id, void *stream);

/* Device views of the live rows, for evaluators that stay on the GPU: row_slot_dev[r] = the slot whose
 * leaf is row r (0xffffffff: a speculative row, nobody waits for it), row_model_dev[r] = the model id that has to evaluate it (tournaments play several
 * models in one batch, rust/src/self_play.rs:203-220).  Valid for r < n_rows after every step(). */
int c4a0_engine_rows_dev(c4a0_engine *e, uint32_t **row_slot_dev, uint64_t **row_model_dev);

/* Device addresses of the row count of the batch to evaluate, for an evaluator that sizes itself on the
 * device (c4a0_net_bind_row_count): the batch has max(*closed_dev, *open_dev) rows.  *closed_dev is set
 * when a tick closes (= c4a0_progress.n_rows); *open_dev is the running counter of a tick that a burst of
 * arena compactions keeps open for k_tail while the evaluator is already enqueued behind k_step. */
int c4a0_engine_rows_count_dev(c4a0_engine *e, const uint32_t **closed_dev, const uint32_t **open_dev);

/* ---- the host loop ------------------------------------------------------------------------
 * A network evaluator as CUDA graphs: graph_exec (a cudaGraphExec_t) reads rows [0, rows) of the
 * engine's planes buffer and writes the same rows of its logits / q buffers.  Pass several sizes,
 * sorted ascending, the largest covering c4a0_engine_io_rows(); every tick the smallest one that covers n_rows is
 * launched. */
typedef struct {
  uint32_t rows;
  void *graph_exec;
} c4a0_nn_graph;

typedef struct {
  uint64_t ticks;            /* tree ticks launched, all engines */
  uint64_t nn_launches;      /* network graphs launched */
  uint64_t nn_rows_launched; /* sum of their row counts (>= nn_evals: bucket padding, re-launches) */
  uint64_t nn_relaunches;    /* ticks whose network was run again because more rows than guessed were live */
  uint64_t tail_launches;    /* ticks that needed the separate compaction kernel */
  double device_ms;          /* CUDA-event time, first network launch .. last game finished (max over engines) */
  double wall_ms;
  double host_wait_ms;       /* host time spent waiting for tick status (the GPU is busy) */
  double host_launch_ms;     /* host time spent enqueuing network graphs and tick kernels */
  uint32_t kernel_samples;   /* ticks whose kernels were bracketed by CUDA events */
  uint32_t reserved;
  double k_step_ms_sum;      /* summed device time of k_step over those ticks */
  double nn_ms_sum;          /* ... of the network graph enqueued behind it */
  uint64_t bucket_launches[32]; /* network launches per graph index (all engines) */
} c4a0_run_report;

/* Plays every engine's requests to completion: what self_play() does between spawning its threads
 * and collecting done_queue (self_play.rs:60-129).  engines[i] runs on streams[i] with graphs[i]
 * (n_graphs[i] entries).  The network of a tick is enqueued right behind the tick's kernel, sized
 * from the previous tick's row count plus a margin, so the host round trip overlaps it; a tick that
 * packed more rows than guessed gets its network run again (nn_relaunches).  time_kernels_every = k > 0 brackets the tree kernels of every k-th tick
 * with CUDA events (no synchronisation).  max_ticks = 0 means no limit. */
int c4a0_engine_run(c4a0_engine *const *engines, uint32_t n_engines,
                    const c4a0_nn_graph *const *graphs, const uint32_t *n_graphs,
                    void *const *streams, uint64_t max_ticks, uint32_t time_kernels_every,
                    c4a0_run_report *out);

/* The same loop with the library's own network kernel as the evaluator (include/c4a0_net.h): nets[i] must read
 * engine i's planes (c4a0_engine_bind_io on c4a0_net_buffer 0), write its logits / q buffers (c4a0_net_bind_outputs)
 * and be bound to its row counters (c4a0_net_bind_row_count(c4a0_engine_rows_count_dev)).  No buckets, no row
 * guess; the two kernels of a tick are launched programmatically dependent on each other, so each one's CTAs are
 * resident and set up while the other drains (set C4A0_PDL=0 in the environment for ordinary launches). */
struct c4a0_net;
int c4a0_engine_run_net(c4a0_engine *const *engines, uint32_t n_engines, struct c4a0_net *const *nets,
                        void *const *streams, uint64_t max_ticks, uint32_t time_kernels_every,
                        c4a0_run_report *out);

/* Finished games [first, first+n) in request order: n_samples[i] (0 = not finished) and
 * [n][43] sample fields (types.rs:104-110).  Any output pointer may be NULL. */
int c4a0_engine_fetch_results(c4a0_engine *e, uint32_t first, uint32_t n, uint32_t *n_samples,
                              uint64_t *mask, uint64_t *value, float *policy, float *q_penalty,
                              float *q_no_penalty, void *stream);
/* Device pointers of the same store ([max_requests][43]...), for zero-copy export to the trainer. */
int c4a0_engine_results_dev(c4a0_engine *e, uint32_t **n_samples_dev, uint64_t **mask_dev,
                            uint64_t **value_dev, float **policy_dev, float **q_penalty_dev,
                            float **q_no_penalty_dev);

/* Training tensors on the device, no host round trip (replaces the per-sample Python loop of
 * src/c4a0/training.py:317-333).  offsets_dev[i] = number of samples of games first..first+i-1 (an
 * exclusive prefix sum of n_samples, computed by the caller), total = their sum.  Sample k of game
 * first+i is written to row offsets_dev[i]+k of pos_dev [rows][2][6][7] f32, policy_dev [rows][7],
 * q_penalty_dev / q_no_penalty_dev [rows]; with flip != 0 its horizontal mirror image (Sample::flip_h,
 * types.rs:115-122) is written to row total+offsets_dev[i]+k as well, so rows = 2*total. */
int c4a0_engine_export_samples(c4a0_engine *e, uint32_t first, uint32_t n, const uint32_t *offsets_dev,
                               uint32_t total, int flip, float *pos_dev, float *policy_dev,
                               float *q_penalty_dev, float *q_no_penalty_dev, void *stream);

/* ---- introspection used by the parity tests -------------------------------------------- */
typedef struct {
  uint32_t state, request, n_moves, root_visits;
  uint64_t root_mask, root_value;
  float root_q_sum_penalty, root_q_sum_no_penalty;
  uint32_t n_blocks; /* expanded nodes currently allocated in the live arena half */
  uint32_t nn_row;   /* network row holding this game's answer (valid in C4A0_ROW_WAIT_NN) */
} c4a0_slot_info;
int c4a0_engine_slot_info(c4a0_engine *e, uint32_t slot, c4a0_slot_info *out, void *stream);
/* Canonical pre-order dump of a slot's tree: 4 words {root kind, N, Qp bits, Qn bits}, then for
 * every expanded node 7 child records of 5 words {kind (0 illegal, 1 leaf, 2 expanded), N, Qp bits,
 * Qn bits, prior bits}, each expanded child followed by its own records.  *needed = words of the
 * full dump; at most `cap` are written. */
int c4a0_engine_dump_tree(c4a0_engine *e, uint32_t slot, uint32_t *buf, size_t cap, size_t *needed,
                          void *stream);

/* ---- stand-alone batch kernels (rules / math / sampling), host buffers in and out --------- */
/* For n positions: terminal state (c4r.rs:228-238), legal-move bitmask (c4r.rs:266-269), ply,
 * terminal values (c4r.rs:253-263), the 7 successor positions (c4r.rs:58-72; zeros when illegal),
 * the NN planes (c4r.rs:378-392) and the mirrored position (c4r.rs:289-299).  NULL outputs skipped. */
int c4a0_rules_batch(int device, const uint64_t *mask, const uint64_t *value, size_t n,
                     float c_ply_penalty, int32_t *terminal, uint32_t *legal, int32_t *ply,
                     float *q_penalty, float *q_no_penalty, uint64_t *child_mask,
                     uint64_t *child_value, float *planes, uint64_t *flip_mask,
                     uint64_t *flip_value);
enum { C4A0_MATH_LOGF = 0, C4A0_MATH_EXPF = 1 };
int c4a0_math_batch(int device, int op, const float *in, float *out, size_t n);
/* softmax over legal moves (c4r.rs:272-286 + mcts.rs:416-434): logits [n][7], legal bitmask [n] */
int c4a0_softmax_batch(int device, const float *logits, const uint32_t *legal, float *out, size_t n);
/* apply_temperature (mcts.rs:439-454) then the seeded weighted draw (mcts.rs:214-222):
 * policy [n][7], temperature [n], seed [n] -> tempered [n][7], column [n] (-1 = reference panics) */
int c4a0_sample_batch(int device, const float *policy, const float *temperature,
                      const uint64_t *seed, float *tempered, int32_t *column, size_t n);

/* Output stage of the network (src/c4a0/nn.py:116-130: log_softmax over the 7 policy outputs, tanh of
 * the two value outputs), straight into the buffers bound with c4a0_engine_bind_io(): one kernel
 * instead of the casts, slices, softmax, tanh and copies PyTorch would launch, which is what a tick
 * with a small batch spends a third of its network time on.  All pointers are DEVICE pointers; the
 * kernel is enqueued on `stream` (capturable into a CUDA graph).
 *   policy_head [rows][ld_policy], value_head [rows][ld_value]: the last linear layers' outputs
 *   (>= 7 and >= 2 valid columns), dtype C4A0_PLANES_F32 or C4A0_PLANES_BF16. */
int c4a0_head_epilogue(const void *policy_head, const void *value_head, uint32_t dtype,
                       uint32_t ld_policy, uint32_t ld_value, uint32_t rows, float *logits_dev,
                       float *q_penalty_dev, float *q_no_penalty_dev, void *stream);

/* The two output layers AND the output stage in one kernel: policy = log_softmax(hp . Wp + bp) over 7
 * columns, values = tanh(hv . Wv + bv) over 2 (src/c4a0/nn.py:100-130), from the heads' last hidden
 * activations hp / hv [rows][ld] (F valid columns, F a multiple of 8; C4A0_PLANES_F32 or _BF16), with
 * the weights given transposed, wp_t [7][F] and wv_t [2][F] in the activations' dtype, biases f32.
 * f32 accumulation on the CUDA cores: one launch instead of two 8-column GEMMs and the output stage,
 * which pays for batches up to a few thousand rows (above that the tensor-core GEMMs win).  All
 * pointers are device pointers; the kernel is enqueued on `stream` (capturable). */
int c4a0_heads(const void *hp, const void *hv, uint32_t dtype, uint32_t ld_hp, uint32_t ld_hv, uint32_t F,
               const void *wp_t, const float *bp, const void *wv_t, const float *bv, uint32_t rows,
               float *logits_dev, float *q_penalty_dev, float *q_no_penalty_dev, void *stream);

/* The same math compiled for the host (no GPU needed); used to pin the restated logf/expf and the
 * sampler against libm / the oracle in the CPU test-suite. */
void c4a0_host_logf(const float *in, float *out, size_t n);
void c4a0_host_expf(const float *in, float *out, size_t n);
int c4a0_host_sample(const float *policy, float temperature, uint64_t seed, float *tempered);
int c4a0_host_terminal_state(uint64_t mask, uint64_t value);
void c4a0_host_make_move(uint64_t mask, uint64_t value, int col, uint64_t *out_mask,
                         uint64_t *out_value);
void c4a0_host_flip_h(uint64_t mask, uint64_t value, uint64_t *out_mask, uint64_t *out_value);
/* the evaluation cache's 49-bit key of a position (injective on positions reachable by play) */
uint64_t c4a0_host_pos_key(uint64_t mask, uint64_t value);
/* idx[0..n) permuted like `results.shuffle(&mut StdRng::seed_from_u64(seed))` (pybridge.rs:110-113) */
void c4a0_host_shuffle(uint64_t seed, uint32_t *idx, size_t n);
/* The pieces of rand 0.10.1's StdRng as restated here, for known-answer tests against public vectors:
 * SeedableRng::seed_from_u64's PCG32 key expansion, and the ChaCha12 word stream of a 256-bit key
 * (64-bit block counter from 0, stream 0): out[0..n) = next_u32() repeatedly. */
void c4a0_host_seed_to_key(uint64_t seed, uint32_t *key8);
void c4a0_host_stdrng_words(const uint32_t *key8, uint32_t *out, size_t n);

/* ---- wire format of PlayGamesResult (host only) ------------------------------------------------------
 * The reference pickles a PlayGamesResult as the CBOR bytes of `serde_cbor::to_vec` (rust/src/pybridge.rs:
 * 73-92, 94-104).  to_cbor: meta [n_games][3] (game_id, player0_id, player1_id), n_samples [n_games] and the
 * padded sample arrays [n_games][43] (policy [n_games][43][7]) -> out; *needed = bytes the encoding takes
 * (call with out = NULL to size the buffer).  from_cbor: call with the arrays NULL to get *n_games, then again
 * with arrays of that many games (cells past a game's sample count are left untouched: pass zeroed arrays).
 * Unknown map keys are skipped; a missing field or malformed input is C4A0_E_INVALID (the reference raises
 * ValueError, pybridge.rs:254-259). */
int c4a0_results_to_cbor(const uint64_t *meta, const uint32_t *n_samples, const uint64_t *mask,
                         const uint64_t *value, const float *policy, const float *q_penalty,
                         const float *q_no_penalty, uint32_t n_games, uint8_t *out, size_t cap, size_t *needed);
int c4a0_results_from_cbor(const uint8_t *buf, size_t len, uint32_t *n_games, uint64_t *meta,
                           uint32_t *n_samples, uint64_t *mask, uint64_t *value, float *policy,
                           float *q_penalty, float *q_no_penalty);

#ifdef __cplusplus
}
#endif
#endif
