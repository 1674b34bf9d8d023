/*
 * c4a0_oracle.h — CPU ORACLE (test infrastructure, NOT product code).
 *
 * A plain-C restatement of the reference's self-play hot path
 * (rust/src/c4r.rs, rust/src/mcts.rs, rust/src/self_play.rs:268-323,
 * rust/src/types.rs:62-161, rust/src/pybridge.rs:110-157).  It exists only so that
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs can check (and time) the CUDA engine against the reference's semantics.
 * Nothing under c4a0_b200/ or c4a0_rust/ may include, link or call it.
 *
 * Pinning: the rules + MCTS semantics are checked against every known-answer
 * test the reference's own test-suite holds for this path (tests/test_oracle_*.py
 * cite them one by one).  PARITY UNPINNED at one boundary only: move sampling
 * and split_train_test use the third-party crate `rand 0.10.1`
 * (rust/Cargo.lock:1585-1591; StdRng = ChaCha12 from `chacha20 0.10.1`), whose
 * source is not under /root/reference and which no reference test pins.  The
 * restatement here follows the crate's published algorithm (see c4o_rng_*).
 *
 * Floating point: all arithmetic is f32 in the reference's operation order,
 * compiled with -ffp-contract=off; ln/exp/sqrt are libm logf/expf/sqrtf, which
 * is what Rust's f32::ln/exp/sqrt lower to (mcts.rs:379, 430, 451-453).
 */
#ifndef C4A0_ORACLE_H
#define C4A0_ORACLE_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define C4O_N_ROWS 6
#define C4O_N_COLS 7
#define C4O_BUF_LEN 84
#define C4O_MAX_SAMPLES 43

/* c4r.rs:13-17 */
typedef struct { uint64_t mask, value; } c4o_pos;

/* c4r.rs:27-32; 0 = not terminal */
enum { C4O_NONE = 0, C4O_PLAYER_WIN = 1, C4O_OPPONENT_WIN = 2, C4O_DRAW = 3 };

/* types.rs:104-110 */
typedef struct {
  c4o_pos pos;
  float policy[C4O_N_COLS];
  float q_penalty, q_no_penalty;
} c4o_sample;

/* types.rs:36-48 */
typedef struct { uint64_t game_id, player0_id, player1_id; } c4o_metadata;

/* ---- rules (c4r.rs) ---- */
int c4o_make_move(c4o_pos p, int col, c4o_pos *out);
int c4o_get(c4o_pos p, int row, int col); /* -1 empty, 0 opponent, 1 player */
int c4o_ply(c4o_pos p);
c4o_pos c4o_invert(c4o_pos p);
int c4o_terminal_state(c4o_pos p);
int c4o_terminal_value(c4o_pos p, float c_ply_penalty, float *qp, float *qn);
unsigned c4o_legal_moves(c4o_pos p); /* bit c set = column c legal */
c4o_pos c4o_flip_h(c4o_pos p);
void c4o_write_planes(c4o_pos p, float *buf84);
const uint64_t *c4o_win_masks(void); /* 69 entries, reference order */
int c4o_from_moves(const int *moves, int n, c4o_pos *out);
int c4o_from_str(const char *utf8, c4o_pos *out);
int c4o_to_str(c4o_pos p, char *buf, size_t cap);
/* proptest strategy random_pos() (c4r.rs:610-629) driven by a caller-supplied column list */
c4o_pos c4o_random_pos(const uint8_t *cols, int n);

/* ---- policy math (mcts.rs:416-454) ---- */
int c4o_softmax(const float in[7], float out[7]); /* 0 on all -inf (reference panics) */
void c4o_apply_temperature(const float in[7], float temperature, float out[7]);

/* ---- rand 0.10.1 restatement (UNPINNED, see header) ---- */
void c4o_seed_from_u64(uint64_t state, uint32_t key[8]);
void c4o_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]);
void c4o_chacha_block_nonce(const uint32_t key[8], uint32_t w12, uint32_t w13, uint32_t w14,
                            uint32_t w15, int rounds, uint32_t out[16]);
int c4o_weighted_index_sample(const float w[7], uint64_t seed); /* -1 on invalid weights */
void c4o_shuffle_indices(uint64_t seed, uint32_t *idx, size_t n);

/* ---- MCTS game (mcts.rs:27-314) ---- */
typedef struct c4o_game c4o_game;
c4o_game *c4o_game_new(c4o_pos pos, c4o_metadata md);
void c4o_game_free(c4o_game *g);
c4o_pos c4o_game_root_pos(const c4o_game *g);
c4o_pos c4o_game_leaf_pos(const c4o_game *g);
uint64_t c4o_game_leaf_model_id(const c4o_game *g);
void c4o_game_on_received_policy(c4o_game *g, const float policy[7], float q_penalty,
                                 float q_no_penalty, float c_exploration, float c_ply_penalty);
int c4o_game_make_move(c4o_game *g, int col, float c_exploration);
int c4o_game_make_random_move(c4o_game *g, float c_exploration, float temperature);
uint64_t c4o_game_root_visit_count(const c4o_game *g);
void c4o_game_root_policy(const c4o_game *g, float out[7]);
float c4o_game_root_q_penalty(const c4o_game *g);
float c4o_game_root_q_no_penalty(const c4o_game *g);
int c4o_game_n_moves(const c4o_game *g);
int c4o_game_to_result(const c4o_game *g, float c_ply_penalty, c4o_sample *out);
/* Canonical pre-order dump of the tree under the current root.  Header: 4 u32 words
 * {root kind, root N, root Qp bits, root Qn bits}; then, for an expanded node, 7 child
 * records of 5 u32 words {kind, N, Qp bits, Qn bits, P bits} (kind: 0 illegal, 1 not
 * expanded, 2 expanded), each expanded child followed recursively by its own 7 records.
 * Returns the number of words the full dump needs (only the first `cap` are written). */
size_t c4o_game_dump_tree(const c4o_game *g, uint32_t *buf, size_t cap);

/* ---- evaluators ---- */
typedef void (*c4o_eval_fn)(void *user, uint64_t model_id, int n, const c4o_pos *pos,
                            float *policy /* n*7 */, float *q_penalty, float *q_no_penalty);
void c4o_eval_uniform(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                      float *qp, float *qn);
/* deterministic integer-hash pseudo network (spec in DESIGN.md §parity, tier E1) */
void c4o_eval_hash(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                   float *qp, float *qn);

void c4o_eval_hash_flat(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                        float *qp, float *qn);

/* ---- per-game state machine of MctsThread::loop_once (self_play.rs:268-323), run
 * game after game on one thread; per-game records do not depend on scheduling. ---- */
typedef struct {
  uint64_t sims;          /* on_received_policy calls (reference-counted) */
  uint64_t nn_evals;      /* leaf positions sent to the evaluator (incl. terminal leaves) */
  uint64_t terminal_leaf_sims;
  uint64_t terminal_root_sims;
  uint64_t moves;
  uint64_t samples;
  uint64_t select_depth_sum; /* sum over sims of the depth of the leaf evaluated */
} c4o_stats;

/* out_samples: n_games * C4O_MAX_SAMPLES; out_n: n_games.  Returns 0 ok. */
int c4o_self_play(const c4o_metadata *reqs, size_t n_games, int max_nn_batch_size,
                  uint64_t n_mcts_iterations, float c_exploration, float c_ply_penalty,
                  c4o_eval_fn eval, void *user, c4o_sample *out_samples, int *out_n,
                  c4o_stats *stats);

/* selfplay_threads.cpp: the reference's thread/channel architecture (timed CPU baseline) */
int c4o_self_play_threaded(const c4o_metadata *reqs, size_t n_games, int max_nn_batch_size,
                           uint64_t n_mcts_iterations, float c_exploration, float c_ply_penalty,
                           c4o_eval_fn eval, void *user, int n_mcts_threads, c4o_sample *out_samples,
                           int *out_n, c4o_stats *stats, uint64_t *nn_batches);
int c4o_self_play_threaded_budget(const c4o_metadata *reqs, size_t n_games, int max_nn_batch_size,
                                  uint64_t n_mcts_iterations, float c_exploration,
                                  float c_ply_penalty, c4o_eval_fn eval, void *user,
                                  int n_mcts_threads, uint64_t max_sims, c4o_sample *out_samples,
                                  int *out_n, c4o_stats *stats, uint64_t *nn_batches);

float c4o_player0_score(const c4o_sample *samples, int n);

/* libm pass-throughs used by tests that pin the product's restated logf/expf */
void c4o_logf_array(const float *in, float *out, size_t n);
void c4o_expf_array(const float *in, float *out, size_t n);

#ifdef __cplusplus
}
#endif
#endif
