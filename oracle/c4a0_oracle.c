/*
 * c4a0_oracle.c — CPU ORACLE (test infrastructure, NOT product code).  See c4a0_oracle.h.
 *
 * Every function cites the reference lines it restates.  The data structures are the
 * reference's own (a heap-allocated pointer tree, one node per position) on purpose: the
 * oracle should share as little as possible with the CUDA engine it checks.
 *
 * Build: gcc -O2 -ffp-contract=off -fPIC -shared (oracle/Makefile).
 */
#include "c4a0_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------
 * Rules — rust/src/c4r.rs
 * ---------------------------------------------------------------------------------------- */

/* c4r.rs:119-122 `_idx_mask_unsafe`: bit = row * 7 + col */
static inline uint64_t idx_mask(int row, int col) { return (uint64_t)1 << (row * C4O_N_COLS + col); }

/* c4r.rs:125-129 `invert` */
c4o_pos c4o_invert(c4o_pos p) {
  p.value = ~p.value;
  p.value &= p.mask;
  return p;
}

/* c4r.rs:58-72 `make_move`: lowest empty row of the column gets a Player stone, then invert.
 * (The reference's bound check is `col > N_COLS`; col == 7 would index bit 7.. of the next row,
 * no caller ever passes it, we reject col >= 7.) */
int c4o_make_move(c4o_pos p, int col, c4o_pos *out) {
  if (col < 0 || col >= C4O_N_COLS) return 0;
  for (int row = 0; row < C4O_N_ROWS; row++) {
    uint64_t idx = idx_mask(row, col);
    if ((idx & p.mask) == 0) {
      p.mask |= idx;
      p.value |= idx;
      *out = c4o_invert(p);
      return 1;
    }
  }
  return 0;
}

/* c4r.rs:76-91 `get` */
int c4o_get(c4o_pos p, int row, int col) {
  if (col < 0 || row < 0 || col >= C4O_N_COLS || row >= C4O_N_ROWS) return -1;
  uint64_t idx = idx_mask(row, col);
  if ((p.mask & idx) == 0) return -1;
  return (p.value & idx) ? 1 : 0;
}

/* c4r.rs:95-97 `ply` */
int c4o_ply(c4o_pos p) { return __builtin_popcountll(p.mask); }

/* c4r.rs:165-224 `WIN_MASKS`: 24 horizontal, 21 vertical, 12 diagonal (row+1,col+1) from rows
 * 0..2, 12 diagonal (row-1,col+1) from rows 3..5 — built in the reference's loop order. */
static uint64_t g_win_masks[69];
static int g_win_masks_ready = 0;
static void build_win_masks(void) {
  int index = 0;
  for (int row = 0; row < C4O_N_ROWS; row++)
    for (int col = 0; col <= C4O_N_COLS - 4; col++)
      g_win_masks[index++] = idx_mask(row, col) | idx_mask(row, col + 1) | idx_mask(row, col + 2) |
                             idx_mask(row, col + 3);
  for (int col = 0; col < C4O_N_COLS; col++)
    for (int row = 0; row <= C4O_N_ROWS - 4; row++)
      g_win_masks[index++] = idx_mask(row, col) | idx_mask(row + 1, col) | idx_mask(row + 2, col) |
                             idx_mask(row + 3, col);
  for (int row = 0; row <= C4O_N_ROWS - 4; row++)
    for (int col = 0; col <= C4O_N_COLS - 4; col++)
      g_win_masks[index++] = idx_mask(row, col) | idx_mask(row + 1, col + 1) |
                             idx_mask(row + 2, col + 2) | idx_mask(row + 3, col + 3);
  for (int row = 3; row < C4O_N_ROWS; row++)
    for (int col = 0; col <= C4O_N_COLS - 4; col++)
      g_win_masks[index++] = idx_mask(row, col) | idx_mask(row - 1, col + 1) |
                             idx_mask(row - 2, col + 2) | idx_mask(row - 3, col + 3);
  if (index != 69) abort();
  g_win_masks_ready = 1;
}
const uint64_t *c4o_win_masks(void) {
  if (!g_win_masks_ready) build_win_masks();
  return g_win_masks;
}

/* c4r.rs:241-249 `_is_terminal_for_player` */
static int is_terminal_for_player(c4o_pos p) {
  const uint64_t *wm = c4o_win_masks();
  uint64_t player_tokens = p.mask & p.value;
  for (int i = 0; i < 69; i++)
    if (__builtin_popcountll(player_tokens & wm[i]) == 4) return 1;
  return 0;
}

/* c4r.rs:228-238 `is_terminal_state`: Player four, then Opponent four, then draw at ply 42 */
int c4o_terminal_state(c4o_pos p) {
  if (is_terminal_for_player(p)) return C4O_PLAYER_WIN;
  if (is_terminal_for_player(c4o_invert(p))) return C4O_OPPONENT_WIN;
  if (c4o_ply(p) == C4O_N_COLS * C4O_N_ROWS) return C4O_DRAW;
  return C4O_NONE;
}

/* c4r.rs:253-263 `terminal_value_with_ply_penalty` */
int c4o_terminal_value(c4o_pos p, float c_ply_penalty, float *qp, float *qn) {
  float ply_penalty_magnitude = c_ply_penalty * (float)c4o_ply(p);
  int t = c4o_terminal_state(p);
  switch (t) {
    case C4O_PLAYER_WIN:
      *qp = 1.0f - ply_penalty_magnitude;
      *qn = 1.0f;
      return t;
    case C4O_OPPONENT_WIN:
      *qp = -1.0f + ply_penalty_magnitude;
      *qn = -1.0f;
      return t;
    case C4O_DRAW:
      *qp = 0.0f;
      *qn = 0.0f;
      return t;
    default:
      return C4O_NONE;
  }
}

/* c4r.rs:266-269 `legal_moves`: top-row cell empty */
unsigned c4o_legal_moves(c4o_pos p) {
  unsigned legal = 0;
  for (int col = 0; col < C4O_N_COLS; col++)
    if (c4o_get(p, C4O_N_ROWS - 1, col) < 0) legal |= 1u << col;
  return legal;
}

/* c4r.rs:289-299 `flip_h` */
c4o_pos c4o_flip_h(c4o_pos p) {
  c4o_pos ret = {0, 0};
  for (int row = 0; row < C4O_N_ROWS; row++)
    for (int col = 0; col < C4O_N_COLS; col++) {
      int piece = c4o_get(p, row, col);
      if (piece >= 0) {
        uint64_t m = idx_mask(row, C4O_N_COLS - 1 - col);
        ret.mask |= m;
        if (piece == 1) ret.value |= m;
      }
    }
  return ret;
}

/* c4r.rs:378-392 `write_numpy_buffer`: [2][6][7] f32, channel 0 = Player, 1 = Opponent */
void c4o_write_planes(c4o_pos p, float *buf) {
  for (int player = 0; player < 2; player++)
    for (int row = 0; row < C4O_N_ROWS; row++)
      for (int col = 0; col < C4O_N_COLS; col++) {
        int idx = player * (C4O_N_ROWS * C4O_N_COLS) + row * C4O_N_COLS + col;
        int cell = c4o_get(p, row, col);
        float v = 0.0f;
        if (cell == 1 && player == 0) v = 1.0f;
        if (cell == 0 && player == 1) v = 1.0f;
        buf[idx] = v;
      }
}

/* c4r.rs:313-319 `from_moves` (the reference unwraps, we return 0 on an illegal move) */
int c4o_from_moves(const int *moves, int n, c4o_pos *out) {
  c4o_pos pos = {0, 0};
  for (int i = 0; i < n; i++)
    if (!c4o_make_move(pos, moves[i], &pos)) return 0;
  *out = pos;
  return 1;
}

/* c4r.rs:415-430 `From<&str>`: lines top row first, red = Player, blue = Opponent.
 * U+1F534 red circle = F0 9F 94 B4, U+1F535 blue circle = F0 9F 94 B5, U+26AB = E2 9A AB. */
int c4o_from_str(const char *s, c4o_pos *out) {
  /* split into lines, then enumerate reversed */
  const char *lines[16];
  size_t lens[16];
  int n_lines = 0;
  const char *p = s;
  while (*p && n_lines < 16) {
    const char *e = strchr(p, '\n');
    size_t len = e ? (size_t)(e - p) : strlen(p);
    lines[n_lines] = p;
    lens[n_lines] = len;
    n_lines++;
    if (!e) break;
    p = e + 1;
  }
  c4o_pos pos = {0, 0};
  for (int li = 0; li < n_lines; li++) {
    int row = n_lines - 1 - li;
    const unsigned char *q = (const unsigned char *)lines[li];
    const unsigned char *end = q + lens[li];
    int col = 0;
    while (q < end) {
      int adv = 1;
      if (*q >= 0xF0) adv = 4; else if (*q >= 0xE0) adv = 3; else if (*q >= 0xC0) adv = 2;
      if (adv == 4 && q + 4 <= end && q[0] == 0xF0 && q[1] == 0x9F && q[2] == 0x94 &&
          (q[3] == 0xB4 || q[3] == 0xB5)) {
        if (row < C4O_N_ROWS && col < C4O_N_COLS) {
          pos.mask |= idx_mask(row, col);
          if (q[3] == 0xB4) pos.value |= idx_mask(row, col);
        }
      }
      q += adv;
      col++;
    }
  }
  *out = pos;
  return 1;
}

/* c4r.rs:395-413 `Display` */
int c4o_to_str(c4o_pos p, char *buf, size_t cap) {
  size_t n = 0;
  for (int row = C4O_N_ROWS - 1; row >= 0; row--) {
    for (int col = 0; col < C4O_N_COLS; col++) {
      int cell = c4o_get(p, row, col);
      const char *g = cell == 1 ? "\xF0\x9F\x94\xB4" : cell == 0 ? "\xF0\x9F\x94\xB5" : "\xE2\x9A\xAB";
      size_t l = strlen(g);
      if (n + l + 2 > cap) return -1;
      memcpy(buf + n, g, l);
      n += l;
    }
    if (row > 0) buf[n++] = '\n';
  }
  buf[n] = 0;
  return (int)n;
}

/* c4r.rs:610-629 proptest strategy `random_pos`: play the listed columns, skipping illegal
 * ones, stopping at the first terminal position. */
c4o_pos c4o_random_pos(const uint8_t *cols, int n) {
  c4o_pos pos = {0, 0};
  for (int i = 0; i < n; i++) {
    if (c4o_terminal_state(pos) != C4O_NONE) break;
    int mov = cols[i] % C4O_N_COLS;
    if (c4o_legal_moves(pos) & (1u << mov)) c4o_make_move(pos, mov, &pos);
  }
  return pos;
}

/* ------------------------------------------------------------------------------------------
 * Policy math — rust/src/mcts.rs:416-454
 * ---------------------------------------------------------------------------------------- */

/* `f32::max` (IEEE maxNum): a NaN operand yields the other one */
static inline float f32_max(float a, float b) {
  if (a != a) return b;
  if (b != b) return a;
  return a > b ? a : b;
}

/* mcts.rs:416-434 `softmax`; returns 0 where the reference panics (all -inf) */
int c4o_softmax(const float in[7], float out[7]) {
  float max = -INFINITY;
  for (int i = 0; i < 7; i++) max = f32_max(max, in[i]);
  if (isinf(max)) return 0;
  float exps[7];
  for (int i = 0; i < 7; i++) exps[i] = expf(in[i] - max);
  float sum = 0.0f; /* `iter().sum::<f32>()` is a left fold */
  for (int i = 0; i < 7; i++) sum = sum + exps[i];
  for (int i = 0; i < 7; i++) out[i] = exps[i] / sum;
  return 1;
}

/* mcts.rs:439-454 `apply_temperature` */
void c4o_apply_temperature(const float in[7], float temperature, float out[7]) {
  int all_eq = 1;
  for (int i = 0; i < 7; i++)
    if (!(in[i] == in[0])) all_eq = 0;
  if (temperature == 1.0f || all_eq) {
    for (int i = 0; i < 7; i++) out[i] = in[i];
    return;
  }
  if (temperature == 0.0f) {
    float max = -INFINITY;
    for (int i = 0; i < 7; i++) max = f32_max(max, in[i]);
    float ret[7], sum = 0.0f;
    for (int i = 0; i < 7; i++) ret[i] = (in[i] == max) ? 1.0f : 0.0f;
    for (int i = 0; i < 7; i++) sum = sum + ret[i];
    for (int i = 0; i < 7; i++) out[i] = ret[i] / sum;
    return;
  }
  float policy_log[7];
  for (int i = 0; i < 7; i++) policy_log[i] = logf(in[i]) / temperature;
  float s = 0.0f;
  for (int i = 0; i < 7; i++) s = s + expf(policy_log[i]);
  float lse = logf(s);
  for (int i = 0; i < 7; i++) {
    float v = expf(policy_log[i] - lse);
    /* f32::clamp(0.0, 1.0): NaN stays NaN */
    if (v < 0.0f) v = 0.0f;
    if (v > 1.0f) v = 1.0f;
    out[i] = v;
  }
}

/* ------------------------------------------------------------------------------------------
 * rand 0.10.1 (third-party, not vendored in /root/reference) — restated from the crate's
 * published algorithm.  PARITY UNPINNED: no reference test holds an expected draw.
 * Call sites: mcts.rs:214-222 (StdRng::seed_from_u64 + WeightedIndex), pybridge.rs:110-113
 * (StdRng + SliceRandom::shuffle).
 * ---------------------------------------------------------------------------------------- */

/* rand_core SeedableRng::seed_from_u64: PCG32 stream fills the 32-byte seed */
void c4o_seed_from_u64(uint64_t state, uint32_t key[8]) {
  const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
  for (int i = 0; i < 8; i++) {
    state = state * MUL + INC;
    uint32_t xorshifted = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xorshifted >> rot) | (xorshifted << ((32 - rot) & 31));
  }
}

static inline uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }
#define QR(a, b, c, d)                                                                            \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12);                           \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

static void chacha_core(const uint32_t init[16], int rounds, uint32_t out[16]) {
  uint32_t x[16];
  memcpy(x, init, sizeof(x));
  for (int i = 0; i < rounds; i += 2) {
    QR(x[0], x[4], x[8], x[12]) QR(x[1], x[5], x[9], x[13])
    QR(x[2], x[6], x[10], x[14]) QR(x[3], x[7], x[11], x[15])
    QR(x[0], x[5], x[10], x[15]) QR(x[1], x[6], x[11], x[12])
    QR(x[2], x[7], x[8], x[13]) QR(x[3], x[4], x[9], x[14])
  }
  for (int i = 0; i < 16; i++) out[i] = x[i] + init[i];
}

/* ChaCha block, 64-bit block counter in words 12-13, stream id 0 in words 14-15 (the layout of
 * rand_chacha / chacha20's ChaCha12Rng).  counter's high half doubles as the first nonce word
 * of the IETF layout, which is how the RFC 7539 vector is checked in the tests. */
void c4o_chacha_block(const uint32_t key[8], uint64_t counter, int rounds, uint32_t out[16]) {
  uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; i++) st[4 + i] = key[i];
  st[12] = (uint32_t)counter;
  st[13] = (uint32_t)(counter >> 32);
  st[14] = 0;
  st[15] = 0;
  chacha_core(st, rounds, out);
}
/* variant with explicit words 14/15, test-only (RFC 7539 §2.3.2 vector) */
void c4o_chacha_block_nonce(const uint32_t key[8], uint32_t w12, uint32_t w13, uint32_t w14,
                            uint32_t w15, int rounds, uint32_t out[16]) {
  uint32_t st[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
  for (int i = 0; i < 8; i++) st[4 + i] = key[i];
  st[12] = w12; st[13] = w13; st[14] = w14; st[15] = w15;
  chacha_core(st, rounds, out);
}

typedef struct {
  uint32_t key[8];
  uint32_t buf[16];
  uint64_t counter;
  int pos;
} std_rng;
static void std_rng_seed(std_rng *r, uint64_t seed) {
  c4o_seed_from_u64(seed, r->key);
  r->counter = 0;
  r->pos = 16;
}
static uint32_t std_rng_next_u32(std_rng *r) {
  if (r->pos >= 16) {
    c4o_chacha_block(r->key, r->counter++, 12, r->buf);
    r->pos = 0;
  }
  return r->buf[r->pos++];
}

/* rand::distr::weighted::WeightedIndex<f32>::new + sample, with
 * rand::distr::uniform::UniformFloat<f32>::new(0, total) as the weight sampler. */
int c4o_weighted_index_sample(const float w[7], uint64_t seed) {
  float cumulative[6];
  float total = w[0];
  if (!(total >= 0.0f)) return -1;
  for (int i = 1; i < 7; i++) {
    if (!(w[i] >= 0.0f)) return -1;
    cumulative[i - 1] = total;
    total = total + w[i];
    if (isinf(total)) return -1;
  }
  if (total == 0.0f) return -1;
  /* UniformFloat::new(low=0, high=total): scale = high - low, shrunk until
   * scale * max_rand + low < high with max_rand = 1 - 2^-23 */
  float low = 0.0f, high = total;
  if (!(low < high) || isinf(high)) return -1;
  float scale = high - low;
  const float max_rand = 1.0f - 1.1920929e-7f;
  while (scale * max_rand + low >= high) {
    uint32_t b;
    memcpy(&b, &scale, 4);
    b -= 1;
    memcpy(&scale, &b, 4);
  }
  std_rng rng;
  std_rng_seed(&rng, seed);
  uint32_t u = std_rng_next_u32(&rng);
  uint32_t fb = (u >> 9) | 0x3f800000u;
  float value1_2;
  memcpy(&value1_2, &fb, 4);
  float value0_1 = value1_2 - 1.0f;
  float x = value0_1 * scale + low;
  /* partition_point(|w| w <= x) over the 6 cumulative weights */
  int idx = 0;
  while (idx < 6 && cumulative[idx] <= x) idx++;
  return idx;
}

/* rand 0.9+/0.10 SliceRandom::shuffle -> partial_shuffle(len) with IncreasingUniform chunks and
 * u32 range sampling by widening multiply with one bias-correction step (Canon's method). */
static uint32_t random_range_u32(std_rng *rng, uint32_t range /* exclusive bound */) {
  /* sample_single_inclusive(0, range-1) */
  if (range == 0) return std_rng_next_u32(rng);
  uint64_t m = (uint64_t)std_rng_next_u32(rng) * range;
  uint32_t result = (uint32_t)(m >> 32), lo_order = (uint32_t)m;
  if (lo_order > (uint32_t)(0u - range)) {
    uint64_t m2 = (uint64_t)std_rng_next_u32(rng) * range;
    uint32_t new_hi = (uint32_t)(m2 >> 32);
    uint32_t sum = lo_order + new_hi;
    if (sum < lo_order) result += 1;
  }
  return result;
}
static void calculate_bound_u32(uint32_t m, uint32_t *bound, uint8_t *count) {
  uint32_t product = m, current = m + 1;
  for (;;) {
    uint64_t p = (uint64_t)product * current;
    if (p <= 0xffffffffULL) {
      product = (uint32_t)p;
      current += 1;
    } else {
      *bound = product;
      *count = (uint8_t)(current - m);
      return;
    }
  }
}
void c4o_shuffle_indices(uint64_t seed, uint32_t *idx, size_t n) {
  if (n <= 1) return;
  std_rng rng;
  std_rng_seed(&rng, seed);
  uint32_t cn = 0, chunk = 0;
  uint8_t chunk_remaining = 1; /* n == 0 -> 1 */
  for (size_t i = 0; i < n; i++) {
    uint32_t next_n = cn + 1;
    uint8_t next_remaining;
    if (chunk_remaining == 0) {
      uint32_t bound;
      uint8_t remaining;
      calculate_bound_u32(next_n, &bound, &remaining);
      chunk = random_range_u32(&rng, bound);
      next_remaining = remaining - 1;
    } else {
      next_remaining = chunk_remaining - 1;
    }
    uint32_t result;
    if (next_remaining == 0) {
      result = chunk;
    } else {
      result = chunk % next_n;
      chunk /= next_n;
    }
    chunk_remaining = next_remaining;
    cn = next_n;
    uint32_t t = idx[i];
    idx[i] = idx[result];
    idx[result] = t;
  }
}

/* ------------------------------------------------------------------------------------------
 * MCTS — rust/src/mcts.rs
 * ---------------------------------------------------------------------------------------- */

/* mcts.rs:332-340 `Node` */
typedef struct node {
  c4o_pos pos;
  struct node *parent; /* Weak<>: NULL once the parent was dropped (i.e. for the root) */
  uint64_t visit_count;
  float q_sum_penalty, q_sum_no_penalty;
  float initial_policy_value;
  int has_children;
  struct node *children[C4O_N_COLS]; /* NULL = illegal move */
} node;

/* mcts.rs:318-322 `RecordedMove` */
typedef struct {
  c4o_pos pos;
  float policy[C4O_N_COLS];
  int mov;
} recorded_move;

/* mcts.rs:27-32 `MctsGame` */
struct c4o_game {
  c4o_metadata metadata;
  node *root;
  node *leaf;
  recorded_move moves[C4O_N_ROWS * C4O_N_COLS + 1];
  int n_moves;
  uint64_t last_select_depth;
};

static node *node_new(c4o_pos pos, node *parent, float prior) { /* mcts.rs:345-355 */
  node *n = (node *)calloc(1, sizeof(node));
  n->pos = pos;
  n->parent = parent;
  n->initial_policy_value = prior;
  return n;
}
static void node_free(node *n) {
  if (!n) return;
  if (n->has_children)
    for (int i = 0; i < C4O_N_COLS; i++) node_free(n->children[i]);
  free(n);
}

c4o_game *c4o_game_new(c4o_pos pos, c4o_metadata md) { /* mcts.rs:48-56: root prior 1.0 */
  c4o_game *g = (c4o_game *)calloc(1, sizeof(c4o_game));
  g->metadata = md;
  g->root = node_new(pos, NULL, 1.0f);
  g->leaf = g->root;
  return g;
}
void c4o_game_free(c4o_game *g) {
  if (!g) return;
  node_free(g->root);
  free(g);
}
c4o_pos c4o_game_root_pos(const c4o_game *g) { return g->root->pos; }
c4o_pos c4o_game_leaf_pos(const c4o_game *g) { return g->leaf->pos; }
int c4o_game_n_moves(const c4o_game *g) { return g->n_moves; }
uint64_t c4o_game_root_visit_count(const c4o_game *g) { return g->root->visit_count; }

/* mcts.rs:70-76 `leaf_model_id_to_play` */
uint64_t c4o_game_leaf_model_id(const c4o_game *g) {
  return (c4o_ply(g->leaf->pos) % 2 == 0) ? g->metadata.player0_id : g->metadata.player1_id;
}

/* mcts.rs:359-367 */
static float node_q_with_penalty(const node *n) {
  return n->q_sum_penalty / ((float)n->visit_count + 1.0f);
}
static float node_q_no_penalty(const node *n) {
  return n->q_sum_no_penalty / ((float)n->visit_count + 1.0f);
}
/* mcts.rs:372-381 `exploration_value` */
static float node_exploration_value(const node *n) {
  float parent_visit_count = n->parent ? (float)n->parent->visit_count : (float)n->visit_count;
  float exploration_value = sqrtf(logf(parent_visit_count) / ((float)n->visit_count + 1.0f));
  return exploration_value * (n->initial_policy_value + 1e-8f);
}
/* mcts.rs:386-388 `uct_value` */
static float node_uct_value(const node *n, float c_exploration) {
  return -node_q_with_penalty(n) + c_exploration * node_exploration_value(n);
}

/* mcts.rs:160-183 `select_new_leaf`: `max_by_key` keeps the LAST maximum */
static void select_new_leaf(c4o_game *g, float c_exploration) {
  node *cur = g->root;
  uint64_t depth = 0;
  while (cur->has_children) {
    node *best = NULL;
    float best_score = 0.0f;
    for (int i = 0; i < C4O_N_COLS; i++) {
      node *child = cur->children[i];
      if (!child) continue;
      float score = node_uct_value(child, c_exploration);
      if (!best || score >= best_score) { /* OrdF32: NaN would panic; never occurs */
        best = child;
        best_score = score;
      }
    }
    if (!best) break;
    cur = best;
    depth++;
  }
  g->leaf = cur;
  g->last_select_depth = depth;
}

/* mcts.rs:114-132 `expand_leaf` */
static void expand_leaf(c4o_game *g, const float policy_probs[7]) {
  node *leaf = g->leaf;
  if (c4o_terminal_state(leaf->pos) != C4O_NONE) return;
  unsigned legal = c4o_legal_moves(leaf->pos);
  for (int m = 0; m < C4O_N_COLS; m++) {
    if (legal & (1u << m)) {
      c4o_pos child_pos;
      c4o_make_move(leaf->pos, m, &child_pos);
      leaf->children[m] = node_new(child_pos, leaf, policy_probs[m]);
    } else {
      leaf->children[m] = NULL;
    }
  }
  leaf->has_children = 1;
}

/* mcts.rs:137-155 `backpropagate_value` */
static void backpropagate_value(c4o_game *g, float q_penalty, float q_no_penalty) {
  node *n = g->leaf;
  for (;;) {
    n->visit_count += 1;
    n->q_sum_penalty += q_penalty;
    n->q_sum_no_penalty += q_no_penalty;
    q_penalty = -q_penalty;
    q_no_penalty = -q_no_penalty;
    if (n->parent) n = n->parent; else break;
  }
}

/* mcts.rs:83-108 `on_received_policy` */
void c4o_game_on_received_policy(c4o_game *g, const float policy_in[7], float q_penalty,
                                 float q_no_penalty, float c_exploration, float c_ply_penalty) {
  c4o_pos leaf_pos = g->leaf->pos;
  float tqp, tqn;
  if (c4o_terminal_value(leaf_pos, c_ply_penalty, &tqp, &tqn) != C4O_NONE) {
    backpropagate_value(g, tqp, tqn);
    select_new_leaf(g, c_exploration);
  } else {
    float logits[7], probs[7];
    unsigned legal = c4o_legal_moves(leaf_pos); /* c4r.rs:272-286 `mask_policy` */
    for (int m = 0; m < 7; m++) logits[m] = (legal & (1u << m)) ? policy_in[m] : -INFINITY;
    if (!c4o_softmax(logits, probs)) abort(); /* reference panics */
    expand_leaf(g, probs);
    backpropagate_value(g, q_penalty, q_no_penalty);
    select_new_leaf(g, c_exploration);
  }
}

/* mcts.rs:396-412 `Node::policy` */
static void node_policy(const node *n, float out[7]) {
  const float uniform = 1.0f / (float)C4O_N_COLS;
  if (!n->has_children) {
    for (int i = 0; i < 7; i++) out[i] = uniform;
    return;
  }
  float counts[7], sum = 0.0f;
  for (int i = 0; i < 7; i++) counts[i] = n->children[i] ? (float)n->children[i]->visit_count : 0.0f;
  for (int i = 0; i < 7; i++) sum = sum + counts[i];
  if (sum == 0.0f) {
    for (int i = 0; i < 7; i++) out[i] = uniform;
  } else {
    for (int i = 0; i < 7; i++) out[i] = counts[i] / sum;
  }
}
void c4o_game_root_policy(const c4o_game *g, float out[7]) { node_policy(g->root, out); }
float c4o_game_root_q_penalty(const c4o_game *g) { return node_q_with_penalty(g->root); }
float c4o_game_root_q_no_penalty(const c4o_game *g) { return node_q_no_penalty(g->root); }

/* mcts.rs:187-206 `make_move`: record (root pos, un-tempered root policy, move); the chosen
 * child becomes the root (its statistics and subtree are kept, everything else is dropped);
 * reselect the leaf.  Returns 0 where the reference panics. */
int c4o_game_make_move(c4o_game *g, int m, float c_exploration) {
  if (m < 0 || m >= C4O_N_COLS) return 0;
  if (!g->root->has_children || !g->root->children[m]) return 0;
  if (g->n_moves >= C4O_N_ROWS * C4O_N_COLS) return 0;
  recorded_move *rm = &g->moves[g->n_moves++];
  rm->pos = g->root->pos;
  node_policy(g->root, rm->policy);
  rm->mov = m;
  node *old_root = g->root;
  node *child = old_root->children[m];
  old_root->children[m] = NULL;
  child->parent = NULL; /* the Weak<> parent dies with the old root */
  node_free(old_root);
  g->root = child;
  select_new_leaf(g, c_exploration);
  return 1;
}

/* mcts.rs:214-222 `make_random_move` */
int c4o_game_make_random_move(c4o_game *g, float c_exploration, float temperature) {
  uint64_t seed = g->metadata.game_id * (uint64_t)((C4O_N_ROWS * C4O_N_COLS) + g->n_moves);
  float policy[7], tempered[7];
  node_policy(g->root, policy);
  c4o_apply_temperature(policy, temperature, tempered);
  int mov = c4o_weighted_index_sample(tempered, seed);
  if (mov < 0) return 0;
  return c4o_game_make_move(g, mov, c_exploration);
}

/* mcts.rs:271-313 `to_result` */
int c4o_game_to_result(const c4o_game *g, float c_ply_penalty, c4o_sample *out) {
  float q_penalty, q_no_penalty;
  if (c4o_terminal_value(g->root->pos, c_ply_penalty, &q_penalty, &q_no_penalty) == C4O_NONE)
    return -1;
  /* cycle [(q, q), (-q, -q)], skipping one when the number of moves is odd */
  int phase = (g->n_moves % 2 == 1) ? 1 : 0;
  int n = 0;
  for (int k = 0; k < g->n_moves; k++, phase ^= 1) {
    out[n].pos = g->moves[k].pos;
    memcpy(out[n].policy, g->moves[k].policy, sizeof(float) * 7);
    out[n].q_penalty = phase ? -q_penalty : q_penalty;
    out[n].q_no_penalty = phase ? -q_no_penalty : q_no_penalty;
    n++;
  }
  out[n].pos = g->root->pos;
  for (int i = 0; i < 7; i++) out[n].policy[i] = 1.0f / (float)C4O_N_COLS;
  out[n].q_penalty = q_penalty;
  out[n].q_no_penalty = q_no_penalty;
  n++;
  return n;
}

static inline uint32_t f32_bits(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  return u;
}
static size_t dump_children(const node *n, uint32_t *buf, size_t cap, size_t w) {
  for (int i = 0; i < C4O_N_COLS; i++) {
    const node *c = n->children[i];
    uint32_t rec[5] = {0, 0, 0, 0, 0};
    if (c) {
      rec[0] = c->has_children ? 2u : 1u;
      rec[1] = (uint32_t)c->visit_count;
      rec[2] = f32_bits(c->q_sum_penalty);
      rec[3] = f32_bits(c->q_sum_no_penalty);
      rec[4] = f32_bits(c->initial_policy_value);
    }
    for (int k = 0; k < 5; k++, w++)
      if (w < cap) buf[w] = rec[k];
    if (c && c->has_children) w = dump_children(c, buf, cap, w);
  }
  return w;
}
size_t c4o_game_dump_tree(const c4o_game *g, uint32_t *buf, size_t cap) {
  size_t w = 0;
  uint32_t head[4] = {g->root->has_children ? 2u : 1u, (uint32_t)g->root->visit_count,
                      f32_bits(g->root->q_sum_penalty), f32_bits(g->root->q_sum_no_penalty)};
  for (int k = 0; k < 4; k++, w++)
    if (w < cap) buf[w] = head[k];
  if (g->root->has_children) w = dump_children(g->root, buf, cap, w);
  return w;
}

/* ------------------------------------------------------------------------------------------
 * Evaluators
 * ---------------------------------------------------------------------------------------- */
void c4o_eval_uniform(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                      float *qp, float *qn) {
  (void)user; (void)model_id; (void)pos;
  for (int i = 0; i < n; i++) {
    for (int k = 0; k < 7; k++) policy[i * 7 + k] = 0.0f;
    qp[i] = 0.0f;
    qn[i] = 0.0f;
  }
}

static inline uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}
/* Tier-E1 pseudo network: h = splitmix64(mask * 0x9E3779B97F4A7C15 ^ splitmix64(value ^ model));
 * logit_k = (16 bits of splitmix64(h + k)) / 8192 - 4   in [-4, 4)
 * q_pen   = (16 bits) / 32768 - 1 scaled by 0.75, q_nopen likewise from other bits.
 * All steps are exact in f32 (integers < 2^24 times powers of two). */
void c4o_eval_hash(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                   float *qp, float *qn) {
  (void)user;
  for (int i = 0; i < n; i++) {
    uint64_t h = splitmix64(pos[i].mask * 0x9E3779B97F4A7C15ULL ^ splitmix64(pos[i].value ^ model_id));
    for (int k = 0; k < 7; k++) {
      uint64_t hk = splitmix64(h + (uint64_t)k);
      policy[i * 7 + k] = (float)(uint32_t)(hk >> 48) * (1.0f / 8192.0f) - 4.0f;
    }
    uint64_t hq = splitmix64(h + 7);
    qp[i] = ((float)(uint32_t)((hq >> 48) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.75f;
    qn[i] = ((float)(uint32_t)((hq >> 32) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.75f;
  }
}

/* The same pseudo network with nearly flat outputs (logits / 16, q scaled by 0.1 instead of 0.75): the
 * shape of a random-init network, so games run as long as gen-0 self-play (~9.5 k simulations per game
 * at 600 per move).  Matches k_eval_builtin(C4A0_EVAL_HASH_FLAT) bit for bit (one f32 multiply each). */
void c4o_eval_hash_flat(void *user, uint64_t model_id, int n, const c4o_pos *pos, float *policy,
                        float *qp, float *qn) {
  (void)user;
  for (int i = 0; i < n; i++) {
    uint64_t h = splitmix64(pos[i].mask * 0x9E3779B97F4A7C15ULL ^ splitmix64(pos[i].value ^ model_id));
    for (int k = 0; k < 7; k++) {
      uint64_t hk = splitmix64(h + (uint64_t)k);
      float lg = (float)(uint32_t)(hk >> 48) * (1.0f / 8192.0f) - 4.0f;
      policy[i * 7 + k] = lg * 0.0625f;
    }
    uint64_t hq = splitmix64(h + 7);
    qp[i] = ((float)(uint32_t)((hq >> 48) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.1f;
    qn[i] = ((float)(uint32_t)((hq >> 32) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.1f;
  }
}

/* ------------------------------------------------------------------------------------------
 * Self-play state machine — rust/src/self_play.rs:268-323 (MctsThread::loop_once) applied to all
 * games in lockstep rounds on one thread.  F8: per-game records do not depend on how
 * games are interleaved, so this serial schedule yields the reference's records.
 * ---------------------------------------------------------------------------------------- */
int c4o_self_play(const c4o_metadata *reqs, size_t n_games, int max_nn_batch_size,
                  uint64_t n_mcts_iterations, float c_exploration, float c_ply_penalty,
                  c4o_eval_fn eval, void *user, c4o_sample *out_samples, int *out_n,
                  c4o_stats *stats) {
  if (max_nn_batch_size <= 0) return -1;
  c4o_stats st;
  memset(&st, 0, sizeof(st));
  c4o_game **games = (c4o_game **)calloc(n_games, sizeof(c4o_game *));
  size_t *pending = (size_t *)malloc(sizeof(size_t) * (n_games ? n_games : 1));
  size_t *batch_ix = (size_t *)malloc(sizeof(size_t) * (size_t)max_nn_batch_size);
  c4o_pos *bpos = (c4o_pos *)malloc(sizeof(c4o_pos) * (size_t)max_nn_batch_size);
  float *bpol = (float *)malloc(sizeof(float) * 7 * (size_t)max_nn_batch_size);
  float *bqp = (float *)malloc(sizeof(float) * (size_t)max_nn_batch_size);
  float *bqn = (float *)malloc(sizeof(float) * (size_t)max_nn_batch_size);
  int rc = 0;
  c4o_pos empty = {0, 0};
  for (size_t i = 0; i < n_games; i++) { /* self_play.rs:55-58 */
    games[i] = c4o_game_new(empty, reqs[i]);
    out_n[i] = 0;
  }
  size_t n_remaining = n_games;
  while (n_remaining > 0 && rc == 0) {
    /* games waiting for an evaluation, in request order */
    size_t n_pending = 0;
    for (size_t i = 0; i < n_games; i++)
      if (games[i]) pending[n_pending++] = i;
    size_t done_mark = 0;
    while (done_mark < n_pending && rc == 0) {
      /* next batch: same model id as the first not-yet-served game, at most max batch */
      size_t first = done_mark;
      while (first < n_pending && pending[first] == (size_t)-1) first++;
      if (first >= n_pending) break;
      uint64_t model_id = c4o_game_leaf_model_id(games[pending[first]]);
      int b = 0;
      for (size_t j = first; j < n_pending && b < max_nn_batch_size; j++) {
        if (pending[j] == (size_t)-1) continue;
        c4o_game *g = games[pending[j]];
        if (c4o_game_leaf_model_id(g) != model_id) continue;
        batch_ix[b] = pending[j];
        bpos[b] = c4o_game_leaf_pos(g);
        pending[j] = (size_t)-1;
        b++;
      }
      done_mark = first;
      eval(user, model_id, b, bpos, bpol, bqp, bqn);
      st.nn_evals += (uint64_t)b;
      for (int k = 0; k < b && rc == 0; k++) {
        size_t gi = batch_ix[k];
        c4o_game *g = games[gi];
        int leaf_terminal = c4o_terminal_state(c4o_game_leaf_pos(g)) != C4O_NONE;
        int root_terminal = c4o_terminal_state(c4o_game_root_pos(g)) != C4O_NONE;
        st.select_depth_sum += g->last_select_depth;
        c4o_game_on_received_policy(g, bpol + 7 * k, bqp[k], bqn[k], c_exploration, c_ply_penalty);
        st.sims++;
        if (root_terminal) st.terminal_root_sims++;
        else if (leaf_terminal) st.terminal_leaf_sims++;
        if (c4o_game_root_visit_count(g) < n_mcts_iterations) continue; /* self_play.rs:283-286 */
        c4o_pos root_pos = c4o_game_root_pos(g);
        if (c4o_terminal_state(root_pos) == C4O_NONE) { /* self_play.rs:290-301 */
          int ply = c4o_ply(root_pos);
          float temperature = ply < 4 ? 4.0f : (ply < 8 ? 2.0f : 1.0f);
          if (!c4o_game_make_random_move(g, c_exploration, temperature)) rc = -2;
          st.moves++;
        } else { /* self_play.rs:302-309 */
          int n = c4o_game_to_result(g, c_ply_penalty, out_samples + gi * C4O_MAX_SAMPLES);
          if (n < 0) rc = -3;
          out_n[gi] = n;
          st.samples += (uint64_t)(n > 0 ? n : 0);
          c4o_game_free(g);
          games[gi] = NULL;
          n_remaining--;
        }
      }
    }
  }
  for (size_t i = 0; i < n_games; i++) c4o_game_free(games[i]);
  free(games); free(pending); free(batch_ix); free(bpos); free(bpol); free(bqp); free(bqn);
  if (stats) *stats = st;
  return rc;
}

/* types.rs:77-99 `player0_score` */
float c4o_player0_score(const c4o_sample *samples, int n) {
  for (int i = 0; i < n; i++) {
    int t = c4o_terminal_state(samples[i].pos);
    if (t != C4O_NONE) {
      float score = t == C4O_PLAYER_WIN ? 1.0f : (t == C4O_OPPONENT_WIN ? 0.0f : 0.5f);
      return (c4o_ply(samples[i].pos) % 2 == 1) ? 1.0f - score : score;
    }
  }
  return NAN; /* reference panics */
}

void c4o_logf_array(const float *in, float *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = logf(in[i]);
}
void c4o_expf_array(const float *in, float *out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = expf(in[i]);
}
