// engine.cu — lockstep MCTS self-play on sm_100a: kernels + the engine half of the C-ABI
// (include/c4a0_engine.h).
//
// What the reference does per game with a heap tree of Rc<RefCell<Node>> (rust/src/mcts.rs:332-355)
// and a thread pool (rust/src/self_play.rs), this file does for thousands of games at once with
// struct-of-arrays trees in HBM:
//
//   * A tree is a bump-allocated array of 160-byte BLOCKS.  A block belongs to one expanded node
//     and holds the statistics of its 7 children column-wise (N[8], Qp[8], Qn[8], P[8], child[8]),
//     i.e. exactly the fields UCT selection reads for one level (mcts.rs:359-388) as five 32-byte
//     sectors.  Node positions are never stored: selection replays the moves on bitboards.
//   * A game is served by 8 lanes (4 games per warp): lane c owns child c, argmax is a 3-step
//     shuffle butterfly with the reference's last-maximum tie-break (Iterator::max_by_key).
//   * Backup (mcts.rs:137-155) walks the path recorded by selection; the path nodes are
//     independent read-modify-writes, so lanes update them in parallel, one f32 add per node per
//     simulation in simulation order — the reference's accumulation order.
//   * A move (mcts.rs:187-222) keeps the chosen child's subtree.  One CTA per moving game copies
//     that subtree breadth-first into the other half of the game's arena (compaction), which is
//     what bounds a game's memory to 2*(n_iterations+2) blocks regardless of game length.
//   * The network batch is dense and de-duplicated, like the reference's NN thread builds it
//     (self_play.rs:203-220: HashSet<Pos> per model): every selected leaf is inserted into an
//     epoch-tagged hash table (smallest slot wins a key), a scan numbers the winners, and only they
//     write input planes, to rows 0..n_rows-1.  Every game remembers the row that holds its answer.
//
// One tick = k_begin -> k_step -> k_move -> k_scan -> k_pack, then the network on rows [0, n_rows).
// No CPU fallback exists: every entry point that computes needs the GPU and fails loudly.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "c4_math.cuh"
#include "c4_rng.cuh"
#include "c4_rules.cuh"
#include "common.cuh"

namespace {

using c4::Pos;
using c4host::blocks_for;
using c4host::fail;

// ------------------------------------------------------------------------------------------------
// Data layout
// ------------------------------------------------------------------------------------------------
struct __align__(32) Block {
  uint32_t N[8];      // visit_count of child c            (mcts.rs:335)
  float Qp[8];        // q_sum_penalty of child c          (mcts.rs:336)
  float Qn[8];        // q_sum_no_penalty of child c       (mcts.rs:337)
  float P[8];         // initial_policy_value of child c   (mcts.rs:338)
  uint32_t child[8];  // block index of child c's own block, 0 = child not expanded
};
static_assert(sizeof(Block) == 160, "block must be five 32-byte sectors");

constexpr int PATH_STRIDE = 44;  // <= 42 levels below a root
constexpr int MAXS = C4A0_MAX_SAMPLES;
constexpr int SCAN_THREADS = 1024;
enum : uint32_t { ST_IDLE = 0, ST_WAIT_NN = 1, ST_CONTINUE = 2, ST_NEED_MOVE = 3 };

struct Globals {  // one instance in device memory
  uint32_t tick;  // epoch of the hash table: leaves selected during tick t carry epoch t
  uint32_t n_req;
  uint32_t next_req;
  uint32_t n_finished;
  uint32_t n_running;
  uint32_t n_movers;
  uint32_t n_rows;
  int32_t error;
  unsigned long long skipped_root_sims;
  unsigned long long moves;
  unsigned long long samples;
  unsigned long long compacted_blocks;
  unsigned long long rows_total;   // sum over ticks of n_rows (positions actually evaluated)
  unsigned long long leaves_total; // sum over ticks of games waiting for the network
};

struct HostStatus {  // mapped pinned host memory, written by k_scan at the end of every tick
  volatile uint32_t tick;
  volatile uint32_t n_rows;
  volatile uint32_t n_finished;
  volatile uint32_t n_running;
  volatile uint32_t n_movers;
  volatile int32_t error;
};

struct Dev {  // passed to kernels by value
  uint32_t n_slots, n_iter, cap, max_inline, plane_bf16, plane_stride, dedup, table_mask;
  float c_expl, c_pen;
  // per slot
  uint64_t *root_mask, *root_value, *leaf_mask, *leaf_value, *leaf_model;
  uint32_t *root_N, *root_block, *half, *n_alloc, *state, *req, *n_moves, *path_len, *path;
  uint32_t *bucket, *urow, *nn_row;
  float *root_Qp, *root_Qn;
  unsigned long long *c_sims, *c_evals, *c_term, *c_depth;
  Block* blocks;  // [n_slots][2][cap]
  // per row
  uint32_t* row_slot;
  // per request
  const uint64_t *game_id, *p0, *p1;
  uint32_t* n_samples;
  uint64_t *s_mask, *s_value;
  float *s_policy, *s_qp, *s_qn;
  // global
  Globals* g;
  HostStatus* status;  // device address of the mapped host struct
  uint32_t* movers;
  unsigned long long* table;  // [table_mask+1] entries: epoch << 32 | leader slot
  // NN io
  void* planes;
  const float *logits, *qp, *qn;
};

__device__ __forceinline__ Block* arena_of(const Dev& D, uint32_t slot, uint32_t half) {
  return D.blocks + ((size_t)slot * 2 + half) * D.cap;
}

__device__ __forceinline__ uint64_t splitmix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ULL;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
  return x ^ (x >> 31);
}

// ------------------------------------------------------------------------------------------------
// 8-lane group helpers
// ------------------------------------------------------------------------------------------------
struct Group {
  unsigned mask;  // the 8 lanes of this game within the warp
  int l;          // 0..7 ; lanes 0..6 own children/columns 0..6
};
__device__ __forceinline__ Group make_group() {
  Group g;
  int lane = threadIdx.x & 31;
  g.mask = 0xffu << (lane & 24);
  g.l = lane & 7;
  return g;
}
template <typename T>
__device__ __forceinline__ T gshfl(const Group& g, T v, int src) {
  return __shfl_sync(g.mask, v, src, 8);
}
template <typename T>
__device__ __forceinline__ T gxor(const Group& g, T v, int m) {
  return __shfl_xor_sync(g.mask, v, m, 8);
}

// NN input planes of `p` into row `row` (c4r.rs:378-392): 84 values, written by 8 lanes as
// 16-byte (f32) or 8-byte (bf16) vectors.
__device__ __forceinline__ void write_planes(const Dev& D, int l, uint32_t row, Pos p) {
  uint64_t mine = p.mask & p.value, theirs = p.mask & ~p.value;
  if (D.plane_bf16) {
    uint2* dst = reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(D.planes) + (size_t)row * D.plane_stride);
    for (int v = l; v < 21; v += 8) {
      uint32_t w[2];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        int i0 = v * 4 + h * 2, i1 = i0 + 1;
        uint32_t b0 = (uint32_t)(((i0 < 42 ? mine >> i0 : theirs >> (i0 - 42))) & 1ull);
        uint32_t b1 = (uint32_t)(((i1 < 42 ? mine >> i1 : theirs >> (i1 - 42))) & 1ull);
        w[h] = (b0 ? 0x3f80u : 0u) | (b1 ? 0x3f800000u : 0u);  // bf16(1.0) = 0x3f80
      }
      dst[v] = make_uint2(w[0], w[1]);
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(D.planes) + (size_t)row * D.plane_stride);
    for (int v = l; v < 21; v += 8) {
      float f[4];
#pragma unroll
      for (int k = 0; k < 4; k++) {
        int i = v * 4 + k;
        f[k] = (float)(((i < 42 ? mine >> i : theirs >> (i - 42))) & 1ull);
      }
      dst[v] = make_float4(f[0], f[1], f[2], f[3]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// Leaf de-duplication (self_play.rs:203-208).  Called by ONE lane per game once its leaf is known.
// Table entry = epoch << 32 | leader slot; an entry whose epoch is not the current tick is empty,
// so the table never needs clearing.  Among games with equal (position, model) the smallest slot
// becomes the leader (atomicMin) — deterministic whatever order the games arrive in.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void publish_leaf(const Dev& D, uint32_t slot, Pos leaf, uint32_t epoch) {
  uint32_t r = D.req[slot];
  uint64_t model = (c4::ply(leaf.mask) & 1) ? D.p1[r] : D.p0[r];  // mcts.rs:70-76
  D.leaf_mask[slot] = leaf.mask;
  D.leaf_value[slot] = leaf.value;
  D.leaf_model[slot] = model;
  if (!D.dedup) return;
  __threadfence();  // the key must be visible before the slot can be found in the table
  uint32_t h = (uint32_t)splitmix64(leaf.mask * 0x9E3779B97F4A7C15ULL ^ splitmix64(leaf.value ^ model)) & D.table_mask;
  const unsigned long long mine = ((unsigned long long)epoch << 32) | slot;
  for (;;) {
    unsigned long long* e = D.table + h;
    unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(e);
    if ((uint32_t)(cur >> 32) != epoch) {
      unsigned long long prev = atomicCAS(e, cur, mine);
      if (prev == cur) break;  // first game with this key in this tick
      cur = prev;
      if ((uint32_t)(cur >> 32) != epoch) continue;
    }
    uint32_t leader = (uint32_t)cur;
    if (__ldcg(D.leaf_mask + leader) == leaf.mask && __ldcg(D.leaf_value + leader) == leaf.value &&
        __ldcg(D.leaf_model + leader) == model) {
      atomicMin(e, mine);
      break;
    }
    h = (h + 1) & D.table_mask;  // another key lives here: linear probing
  }
  D.bucket[slot] = h;
}

// ------------------------------------------------------------------------------------------------
// Per-game working set held in registers by all 8 lanes (group-uniform values)
// ------------------------------------------------------------------------------------------------
struct Game {
  uint32_t slot;
  Pos root;
  uint32_t rootN, root_block, n_alloc, len;
  float rootQp, rootQn;
  Block* arena;
  uint32_t* path;
  unsigned long long sims, evals, term, depth;
};

__device__ __forceinline__ void load_game(const Dev& D, uint32_t slot, Game& G) {
  G.slot = slot;
  G.root.mask = D.root_mask[slot];
  G.root.value = D.root_value[slot];
  G.rootN = D.root_N[slot];
  G.rootQp = D.root_Qp[slot];
  G.rootQn = D.root_Qn[slot];
  G.root_block = D.root_block[slot];
  G.n_alloc = D.n_alloc[slot];
  G.len = D.path_len[slot];
  G.arena = arena_of(D, slot, D.half[slot]);
  G.path = D.path + (size_t)slot * PATH_STRIDE;
  G.sims = G.evals = G.term = G.depth = 0;
}
__device__ __forceinline__ void store_game(const Dev& D, const Group& g, const Game& G, uint32_t state) {
  if (g.l == 0) {
    uint32_t s = G.slot;
    D.root_N[s] = G.rootN;
    D.root_Qp[s] = G.rootQp;
    D.root_Qn[s] = G.rootQn;
    D.root_block[s] = G.root_block;
    D.n_alloc[s] = G.n_alloc;
    D.path_len[s] = G.len;
    D.state[s] = state;
    if (G.sims) D.c_sims[s] += G.sims;
    if (G.evals) D.c_evals[s] += G.evals;
    if (G.term) D.c_term[s] += G.term;
    if (G.depth) D.c_depth[s] += G.depth;
  }
}

// mcts.rs:137-155 — add (qp, qn) at the leaf, alternate the sign towards the root.
__device__ __forceinline__ void backup(const Group& g, Game& G, float qp, float qn) {
  __syncwarp(g.mask);  // path[] written by lane 0 of this group
  for (uint32_t j = g.l; j < G.len; j += 8) {
    uint32_t e = G.path[j];
    Block* B = G.arena + (e >> 3);
    uint32_t c = e & 7u;
    bool neg = ((G.len - 1 - j) & 1u) != 0;
    B->N[c] += 1u;
    B->Qp[c] += neg ? -qp : qp;
    B->Qn[c] += neg ? -qn : qn;
  }
  bool neg = (G.len & 1u) != 0;
  G.rootN += 1u;
  G.rootQp += neg ? -qp : qp;
  G.rootQn += neg ? -qn : qn;
  __syncwarp(g.mask);  // statistics visible to the selection that follows
}

// mcts.rs:160-183 + 359-388 — walk from the root to a node without children.
__device__ __forceinline__ Pos select_leaf(const Dev& D, const Group& g, Game& G) {
  Pos pos = G.root;
  uint32_t np = G.rootN, b = G.root_block, len = 0;
  const float ninf = -c4::f32_inf();
  while (b != 0u && len < 43u) {  // a tree is at most 42 plies deep
    const Block* B = G.arena + b;
    uint32_t n = B->N[g.l];
    float qs = B->Qp[g.l];
    float pr = B->P[g.l];
    uint32_t ch = B->child[g.l];
    unsigned legal = c4::legal_mask(pos.mask);
    float lnp = c4::c4_logf((float)np);             // ln(parent visits), mcts.rs:378-379
    float nf = (float)n + 1.0f;
    float q = qs / nf;                              // mcts.rs:359-361
    float ex = sqrtf(lnp / nf);
    ex = ex * (pr + 1e-8f);                         // mcts.rs:380
    float u = (-q) + (D.c_expl * ex);               // mcts.rs:386-388
    bool ok = (g.l < 7) && ((legal >> g.l) & 1u);
    float bu = ok ? u : ninf;
    int bl = ok ? g.l : -1;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) {               // max_by_key: the LAST maximum wins
      float ou = gxor(g, bu, m);
      int ol = gxor(g, bl, m);
      bool take = (ol >= 0) && (bl < 0 || ou > bu || (ou == bu && ol > bl));
      bu = take ? ou : bu;
      bl = take ? ol : bl;
    }
    if (g.l == 0) G.path[len] = (b << 3) | (uint32_t)bl;
    len++;
    pos = c4::make_move(pos, bl);
    np = gshfl(g, n, bl);
    b = gshfl(g, ch, bl);
  }
  G.len = len;
  return pos;
}

// Run simulations from "leaf unknown" until a leaf needs the network, the root needs to move, or
// the per-step budget of in-kernel (terminal-leaf) simulations is spent.
__device__ __forceinline__ uint32_t advance(const Dev& D, const Group& g, Game& G, uint32_t epoch) {
  for (uint32_t it = 0;; it++) {
    Pos leaf = select_leaf(D, g, G);
    float tqp, tqn;
    int t = c4::terminal_value(leaf, D.c_pen, &tqp, &tqn);
    if (t == c4::NONE) {
      if (g.l == 0) publish_leaf(D, G.slot, leaf, epoch);
      return ST_WAIT_NN;
    }
    // terminal leaf: mcts.rs:92-98 — no expansion, back up the objective value
    G.depth += G.len;
    backup(g, G, tqp, tqn);
    G.sims++;
    G.term++;
    if (G.rootN >= D.n_iter) return ST_NEED_MOVE;   // self_play.rs:283
    if (it + 1 >= D.max_inline) return ST_CONTINUE;
  }
}

__device__ __forceinline__ void push_mover(const Dev& D, uint32_t slot) {
  uint32_t i = atomicAdd(&D.g->n_movers, 1u);
  D.movers[i] = slot;
}

// ------------------------------------------------------------------------------------------------
// K_step: apply network outputs (expand + backup), then select the next leaf.  8 lanes per game.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_step(Dev D) {
  uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (slot >= D.n_slots) return;
  Group g = make_group();
  uint32_t st = D.state[slot];
  if (st == ST_IDLE) return;
  if (st == ST_NEED_MOVE) {  // reached n_iterations inside k_move's own advance()
    if (g.l == 0) push_mover(D, slot);
    return;
  }
  const uint32_t epoch = D.g->tick;
  Game G;
  load_game(D, slot, G);
  if (st == ST_WAIT_NN) {
    // mask_policy + softmax (c4r.rs:272-286, mcts.rs:416-434) over the leaf's legal moves
    const uint32_t row = D.nn_row[slot];
    unsigned legal = c4::legal_mask(D.leaf_mask[slot]);
    bool ok = (g.l < 7) && ((legal >> g.l) & 1u);
    float x = ok ? D.logits[(size_t)row * 7 + g.l] : -c4::f32_inf();
    float mx = x;
#pragma unroll
    for (int m = 1; m < 8; m <<= 1) mx = fmaxf(mx, gxor(g, mx, m));
    float e = ok ? c4::c4_expf(x - mx) : 0.0f;
    float s = 0.0f;
#pragma unroll
    for (int i = 0; i < 7; i++) s = s + gshfl(g, e, i);  // left fold, as iter().sum()
    float p = ok ? e / s : 0.0f;
    // expand_leaf (mcts.rs:114-132): one new block holding the 7 children
    uint32_t nb = G.n_alloc++;
    if (nb >= D.cap) {
      if (g.l == 0) D.g->error = C4A0_E_ENGINE;
      return;
    }
    Block* B = G.arena + nb;
    B->N[g.l] = 0u;
    B->Qp[g.l] = 0.0f;
    B->Qn[g.l] = 0.0f;
    B->P[g.l] = p;
    B->child[g.l] = 0u;
    if (G.len == 0) {
      G.root_block = nb;
    } else if (g.l == 0) {
      uint32_t e2 = G.path[G.len - 1];
      G.arena[e2 >> 3].child[e2 & 7u] = nb;
    }
    G.depth += G.len;
    backup(g, G, D.qp[row], D.qn[row]);
    G.sims++;
    G.evals++;
    if (G.rootN >= D.n_iter) {  // self_play.rs:283-300: time to move (or finish)
      store_game(D, g, G, ST_NEED_MOVE);
      if (g.l == 0) push_mover(D, slot);
      return;
    }
  }
  uint32_t ns = advance(D, g, G, epoch);
  store_game(D, g, G, ns);
  if (ns == ST_NEED_MOVE && g.l == 0) push_mover(D, slot);
}

// ------------------------------------------------------------------------------------------------
// K_move: one CTA per game that has to move.
// ------------------------------------------------------------------------------------------------
constexpr int MOVE_THREADS = 128;

__device__ void seat_game(const Dev& D, uint32_t slot, uint32_t r) {
  D.root_mask[slot] = 0ull;
  D.root_value[slot] = 0ull;
  D.root_N[slot] = 0u;
  D.root_Qp[slot] = 0.0f;
  D.root_Qn[slot] = 0.0f;
  D.root_block[slot] = 0u;
  D.half[slot] = 0u;
  D.n_alloc[slot] = 1u;
  D.req[slot] = r;
  D.n_moves[slot] = 0u;
  D.path_len[slot] = 0u;
  D.state[slot] = ST_WAIT_NN;
}

__global__ void __launch_bounds__(MOVE_THREADS) k_move(Dev D) {
  __shared__ uint32_t sh_src, sh_next, sh_head, sh_mode;  // mode: 0 re-root, 1 game over, 2 error
  __shared__ Pos sh_newpos;
  __shared__ uint32_t sh_newN;
  __shared__ float sh_newQp, sh_newQn, sh_tqp, sh_tqn;
  const uint32_t n_movers = D.g->n_movers;
  const uint32_t epoch = D.g->tick;
  const uint32_t n_req = D.g->n_req;
  for (uint32_t m = blockIdx.x; m < n_movers; m += gridDim.x) {
    const uint32_t slot = D.movers[m];
    const uint32_t req = D.req[slot];
    const uint32_t half = D.half[slot];
    Block* src = arena_of(D, slot, half);
    Block* dst = arena_of(D, slot, half ^ 1u);
    if (threadIdx.x == 0) {
      Pos root{D.root_mask[slot], D.root_value[slot]};
      uint32_t rb = D.root_block[slot];
      uint32_t nm = D.n_moves[slot];
      // root_policy (mcts.rs:396-412): visit counts of the children, normalised
      float pol[7], tempered[7];
      const float uniform = 1.0f / 7.0f;
      if (rb) {
        float cnt[7], sum = 0.0f;
        for (int i = 0; i < 7; i++) cnt[i] = (float)src[rb].N[i];
        for (int i = 0; i < 7; i++) sum = sum + cnt[i];
        for (int i = 0; i < 7; i++) pol[i] = (sum == 0.0f) ? uniform : cnt[i] / sum;
      } else {
        for (int i = 0; i < 7; i++) pol[i] = uniform;
      }
      // self_play.rs:294-299 + mcts.rs:214-222
      float T = c4::temperature_for_ply(c4::ply(root.mask));
      c4::apply_temperature7(pol, T, tempered);
      int col = c4::weighted_sample7(tempered, c4::move_seed(D.game_id[req], (int)nm));
      unsigned legal = c4::legal_mask(root.mask);
      if (col < 0 || !((legal >> col) & 1u) || rb == 0u || nm >= 42u) {
        // the reference panics here (mcts.rs:190-197); park the game and report
        D.g->error = C4A0_E_ENGINE;
        D.state[slot] = ST_IDLE;
        atomicSub(&D.g->n_running, 1u);
        sh_mode = 2u;
      } else {
        // RecordedMove (mcts.rs:198-203) goes straight into the sample store
        size_t si = (size_t)req * MAXS + nm;
        D.s_mask[si] = root.mask;
        D.s_value[si] = root.value;
        for (int i = 0; i < 7; i++) D.s_policy[si * 7 + i] = pol[i];
        D.n_moves[slot] = nm + 1u;
        Pos np = c4::make_move(root, col);
        sh_newpos = np;
        sh_newN = src[rb].N[col];
        sh_newQp = src[rb].Qp[col];
        sh_newQn = src[rb].Qn[col];
        sh_src = src[rb].child[col];
        float tqp, tqn;
        int t = c4::terminal_value(np, D.c_pen, &tqp, &tqn);
        sh_tqp = tqp;
        sh_tqn = tqn;
        sh_mode = (t != c4::NONE) ? 1u : 0u;
        atomicAdd(&D.g->moves, 1ull);
      }
    }
    __syncthreads();
    const uint32_t mode = sh_mode;
    if (mode == 1u) {
      // to_result (mcts.rs:271-313): alternate the terminal value back through the moves
      const uint32_t L = D.n_moves[slot];
      const float tqp = sh_tqp, tqn = sh_tqn;
      for (uint32_t k = threadIdx.x; k < L; k += blockDim.x) {
        bool neg = ((L - k) & 1u) != 0;
        size_t si = (size_t)req * MAXS + k;
        D.s_qp[si] = neg ? -tqp : tqp;
        D.s_qn[si] = neg ? -tqn : tqn;
      }
      if (threadIdx.x == 0) {
        size_t si = (size_t)req * MAXS + L;
        D.s_mask[si] = sh_newpos.mask;
        D.s_value[si] = sh_newpos.value;
        for (int i = 0; i < 7; i++) D.s_policy[si * 7 + i] = 1.0f / 7.0f;
        D.s_qp[si] = tqp;
        D.s_qn[si] = tqn;
        D.n_samples[req] = L + 1u;
        atomicAdd(&D.g->samples, (unsigned long long)(L + 1u));
        atomicAdd(&D.g->n_finished, 1u);
        // the reference keeps simulating the terminal root until N >= n
        uint32_t nN = sh_newN;
        if (nN < D.n_iter) atomicAdd(&D.g->skipped_root_sims, (unsigned long long)(D.n_iter - nN));
        // seat the next waiting request in this slot (self_play.rs:55-58 queues them all up front)
        uint32_t r = atomicAdd(&D.g->next_req, 1u);
        if (r < n_req) {
          seat_game(D, slot, r);
          publish_leaf(D, slot, Pos{0ull, 0ull}, epoch);
        } else {
          D.state[slot] = ST_IDLE;
          atomicSub(&D.g->n_running, 1u);
        }
      }
      __syncthreads();
      continue;
    }
    if (mode == 2u) {
      __syncthreads();
      continue;
    }
    // ---- re-root: copy the kept subtree breadth-first into the other arena half -------------
    if (threadIdx.x == 0) {
      sh_head = 1u;
      sh_next = sh_src ? 2u : 1u;
    }
    if (sh_src && threadIdx.x < 10) {
      reinterpret_cast<uint4*>(dst + 1)[threadIdx.x] = reinterpret_cast<const uint4*>(src + sh_src)[threadIdx.x];
    }
    __syncthreads();
    for (;;) {
      const uint32_t head = sh_head, tail = sh_next;  // blocks [head, tail) form one tree level
      __syncthreads();
      if (head == tail) break;
      if (tail > D.cap) {  // cannot happen for a well-formed tree; never run off the arena
        if (threadIdx.x == 0) D.g->error = C4A0_E_ENGINE;
        break;
      }
      const uint32_t nwork = (tail - head) * 8u;
      for (uint32_t w = threadIdx.x; w < nwork; w += blockDim.x) {
        uint32_t b = head + (w >> 3), c = w & 7u;
        if (c == 7u) continue;
        uint32_t s = dst[b].child[c];  // still an index into src
        if (s) {
          uint32_t j = atomicAdd(&sh_next, 1u);
          const uint4* from = reinterpret_cast<const uint4*>(src + s);
          uint4* to = reinterpret_cast<uint4*>(dst + j);
#pragma unroll
          for (int q = 0; q < 10; q++) to[q] = from[q];
          dst[b].child[c] = j;
        }
      }
      __syncthreads();
      if (threadIdx.x == 0) sh_head = tail;
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      D.root_mask[slot] = sh_newpos.mask;
      D.root_value[slot] = sh_newpos.value;
      D.root_N[slot] = sh_newN;
      D.root_Qp[slot] = sh_newQp;
      D.root_Qn[slot] = sh_newQn;
      D.root_block[slot] = sh_src ? 1u : 0u;
      D.half[slot] = half ^ 1u;
      D.n_alloc[slot] = sh_next;
      D.path_len[slot] = 0u;
      atomicAdd(&D.g->compacted_blocks, (unsigned long long)(sh_next - 1u));
    }
    __syncthreads();
    // mcts.rs:205: select the new leaf under the new root
    if (threadIdx.x < 8) {
      Group g = make_group();
      Game G;
      load_game(D, slot, G);
      uint32_t ns = advance(D, g, G, epoch);
      store_game(D, g, G, ns);
    }
    __syncthreads();
  }
}

// First kernel of a tick: new epoch, empty mover list.
__global__ void k_begin_step(Dev D) {
  D.g->tick += 1u;
  D.g->n_movers = 0u;
}

// Seat the first min(n_slots, n_req) games (self_play.rs:55-58).
__global__ void k_init(Dev D, uint32_t n_req) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot == 0) {
    Globals z;
    memset(&z, 0, sizeof(z));
    z.tick = 1u;
    z.n_req = n_req;
    z.next_req = n_req < D.n_slots ? n_req : D.n_slots;
    z.n_running = z.next_req;
    *D.g = z;
  }
  if (slot >= D.n_slots) return;
  D.c_sims[slot] = D.c_evals[slot] = D.c_term[slot] = D.c_depth[slot] = 0ull;
  if (slot < n_req) {
    seat_game(D, slot, slot);
    publish_leaf(D, slot, Pos{0ull, 0ull}, 1u);
  } else {
    D.state[slot] = ST_IDLE;
  }
}

// ------------------------------------------------------------------------------------------------
// K_scan (one CTA): number the leaders in slot order -> urow[], n_rows; publish the tick's status
// to the host.  K_pack: leaders write their planes to their row, every waiting game records the row
// that will hold its answer.
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool is_leader(const Dev& D, uint32_t slot) {
  if (D.state[slot] != ST_WAIT_NN) return false;
  if (!D.dedup) return true;
  return (uint32_t)D.table[D.bucket[slot]] == slot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan(Dev D) {
  __shared__ uint32_t warp_sums[SCAN_THREADS / 32];
  __shared__ uint32_t warp_wait[SCAN_THREADS / 32];
  const uint32_t per = (D.n_slots + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t lo = threadIdx.x * per;
  const uint32_t hi = min(lo + per, D.n_slots);
  uint32_t cnt = 0, waiting = 0;
  for (uint32_t s = lo; s < hi; s++) {
    cnt += is_leader(D, s) ? 1u : 0u;
    waiting += (D.state[s] == ST_WAIT_NN) ? 1u : 0u;
  }
  // block-wide exclusive scan of cnt
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = cnt;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += v;
  }
  uint32_t wsum = waiting;
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, d);
  if (lane == 31) warp_sums[warp] = incl;
  if (lane == 0) warp_wait[warp] = wsum;
  __syncthreads();
  if (warp == 0) {
    uint32_t v = warp_sums[lane];
    uint32_t iv = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t t = __shfl_up_sync(0xffffffffu, iv, d);
      if (lane >= d) iv += t;
    }
    warp_sums[lane] = iv - v;  // exclusive prefix of the warp totals
    uint32_t ww = warp_wait[lane];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ww += __shfl_xor_sync(0xffffffffu, ww, d);
    if (lane == 31) {
      uint32_t total = iv;
      Globals* G = D.g;
      G->n_rows = total;
      G->rows_total += total;
      G->leaves_total += ww;
      HostStatus* hs = D.status;
      hs->n_rows = total;
      hs->n_finished = G->n_finished;
      hs->n_running = G->n_running;
      hs->n_movers = G->n_movers;
      hs->error = G->error;
      __threadfence_system();
      hs->tick = G->tick;  // written last: the host spins on it
    }
  }
  __syncthreads();
  uint32_t base = warp_sums[warp] + (incl - cnt);
  for (uint32_t s = lo; s < hi; s++)
    if (is_leader(D, s)) D.urow[s] = base++;
}

__global__ void __launch_bounds__(256) k_pack(Dev D) {
  uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  if (slot >= D.n_slots) return;
  if (D.state[slot] != ST_WAIT_NN) return;
  const int l = threadIdx.x & 7;
  uint32_t leader = D.dedup ? (uint32_t)D.table[D.bucket[slot]] : slot;
  uint32_t row = D.urow[leader];
  if (l == 0) D.nn_row[slot] = row;
  if (leader == slot) {
    if (l == 0) D.row_slot[row] = slot;
    write_planes(D, l, row, Pos{D.leaf_mask[slot], D.leaf_value[slot]});
  }
}

// ------------------------------------------------------------------------------------------------
// Synthetic evaluators (parity tiers E0 / E1) and row views
// ------------------------------------------------------------------------------------------------
__global__ void k_eval_builtin(Dev D, int kind, float* logits, float* qp, float* qn) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= D.g->n_rows) return;
  if (kind == C4A0_EVAL_UNIFORM) {
    for (int k = 0; k < 7; k++) logits[(size_t)row * 7 + k] = 0.0f;
    qp[row] = 0.0f;
    qn[row] = 0.0f;
    return;
  }
  uint32_t slot = D.row_slot[row];
  uint64_t mask = D.leaf_mask[slot], value = D.leaf_value[slot], model = D.leaf_model[slot];
  uint64_t h = splitmix64(mask * 0x9E3779B97F4A7C15ULL ^ splitmix64(value ^ model));
  for (int k = 0; k < 7; k++) {
    uint64_t hk = splitmix64(h + (uint64_t)k);
    logits[(size_t)row * 7 + k] = (float)(uint32_t)(hk >> 48) * (1.0f / 8192.0f) - 4.0f;
  }
  uint64_t hq = splitmix64(h + 7);
  qp[row] = ((float)(uint32_t)((hq >> 48) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.75f;
  qn[row] = ((float)(uint32_t)((hq >> 32) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * 0.75f;
}

__global__ void k_gather_rows(Dev D, uint64_t* mask, uint64_t* value, uint64_t* model) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= D.g->n_rows) return;
  uint32_t slot = D.row_slot[row];
  mask[row] = D.leaf_mask[slot];
  value[row] = D.leaf_value[slot];
  model[row] = D.leaf_model[slot];
}

__global__ void k_sum_counters(Dev D, unsigned long long* out4) {
  unsigned long long a = 0, b = 0, c = 0, d = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.n_slots; i += gridDim.x * blockDim.x) {
    a += D.c_sims[i];
    b += D.c_evals[i];
    c += D.c_term[i];
    d += D.c_depth[i];
  }
  atomicAdd(out4 + 0, a);
  atomicAdd(out4 + 1, b);
  atomicAdd(out4 + 2, c);
  atomicAdd(out4 + 3, d);
}

}  // namespace

// ------------------------------------------------------------------------------------------------
// Engine object
// ------------------------------------------------------------------------------------------------
struct c4a0_engine {
  c4a0_config cfg;
  Dev D;
  std::vector<void*> allocs;
  size_t bytes = 0;
  bool io_bound = false, have_requests = false;
  uint32_t n_req = 0;
  uint64_t steps = 0;
  unsigned long long* scratch4 = nullptr;
  uint64_t *row_mask = nullptr, *row_value = nullptr, *row_model = nullptr;
  Globals* h_globals = nullptr;   // pinned
  HostStatus* h_status = nullptr; // pinned + mapped
  float *b_logits = nullptr, *b_qp = nullptr, *b_qn = nullptr;  // writable aliases for eval_builtin
  cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
};

namespace {
template <typename T>
int dalloc(c4a0_engine* e, T** p, size_t n) {
  size_t b = (n ? n : 1) * sizeof(T);
  cudaError_t err = cudaMalloc((void**)p, b);
  if (err != cudaSuccess) {
    cudaGetLastError();
    return fail(err == cudaErrorMemoryAllocation ? C4A0_E_NOMEM : C4A0_E_CUDA,
                "cudaMalloc(%zu bytes) failed: %s", b, cudaGetErrorString(err));
  }
  e->allocs.push_back(*p);
  e->bytes += b;
  return 0;
}
#define DA(ptr, n)                         \
  do {                                     \
    int _r = dalloc(e, &(ptr), (n));       \
    if (_r) {                              \
      c4a0_engine_destroy(e);              \
      return _r;                           \
    }                                      \
  } while (0)

// Enqueue one tick.  `ev` (3 events) brackets k_step and k_move when given.
int launch_tick(c4a0_engine* e, cudaStream_t s, cudaEvent_t* ev) {
  const Dev& D = e->D;
  k_begin_step<<<1, 1, 0, s>>>(D);
  if (ev) CK(cudaEventRecord(ev[0], s));
  k_step<<<blocks_for((size_t)D.n_slots * 8, 256), 256, 0, s>>>(D);
  if (ev) CK(cudaEventRecord(ev[1], s));
  unsigned grid = D.n_slots < 148u * 8u ? D.n_slots : 148u * 8u;
  k_move<<<grid, MOVE_THREADS, 0, s>>>(D);
  if (ev) CK(cudaEventRecord(ev[2], s));
  k_scan<<<1, SCAN_THREADS, 0, s>>>(D);
  k_pack<<<blocks_for((size_t)D.n_slots * 8, 256), 256, 0, s>>>(D);
  CK(cudaGetLastError());
  e->steps++;
  return 0;
}

const char* kEngineErr =
    "a game reached a state where the reference panics (illegal sampled move or arena overflow)";
}  // namespace

extern "C" {

int c4a0_engine_create(const c4a0_config* cfg, c4a0_engine** out) {
  if (!cfg || !out) return fail(C4A0_E_INVALID, "null argument");
  *out = nullptr;
  if (cfg->n_slots == 0 || cfg->n_mcts_iterations == 0 || cfg->max_requests == 0)
    return fail(C4A0_E_INVALID, "n_slots, max_requests and n_mcts_iterations must be >= 1");
  if (cfg->plane_dtype > C4A0_PLANES_BF16) return fail(C4A0_E_INVALID, "bad plane_dtype");
  if (cfg->plane_stride && (cfg->plane_stride < 84 || cfg->plane_stride % 4))
    return fail(C4A0_E_INVALID, "plane_stride must be 0 or a multiple of 4 that is >= 84");
  if (cfg->n_mcts_iterations > (1u << 24))
    return fail(C4A0_E_INVALID, "n_mcts_iterations above 2^24 is not exactly representable in f32");
  if (cfg->n_slots > (1u << 24)) return fail(C4A0_E_INVALID, "n_slots above 2^24 is not supported");
  int r = c4host::no_gpu_error();
  if (r) return r;
  CK(cudaSetDevice(cfg->device));
  c4a0_engine* e = new c4a0_engine();
  e->cfg = *cfg;
  Dev& D = e->D;
  memset(&D, 0, sizeof(D));
  D.n_slots = cfg->n_slots;
  D.n_iter = cfg->n_mcts_iterations;
  D.cap = cfg->n_mcts_iterations + 2;  // index 0 unused; at most n_iter expanded nodes per tree
  D.max_inline = cfg->max_inline_sims ? cfg->max_inline_sims : 8;
  D.plane_bf16 = cfg->plane_dtype == C4A0_PLANES_BF16;
  D.plane_stride = cfg->plane_stride ? cfg->plane_stride : 84;
  D.dedup = (cfg->flags & C4A0_FLAG_NO_DEDUP) ? 0u : 1u;
  D.c_expl = cfg->c_exploration;
  D.c_pen = cfg->c_ply_penalty;
  size_t S = cfg->n_slots, R = cfg->max_requests;
  size_t T = 1;
  while (T < 2 * S) T <<= 1;
  D.table_mask = (uint32_t)(T - 1);
  DA(D.root_mask, S); DA(D.root_value, S); DA(D.leaf_mask, S); DA(D.leaf_value, S); DA(D.leaf_model, S);
  DA(D.root_N, S); DA(D.root_block, S); DA(D.half, S); DA(D.n_alloc, S); DA(D.state, S);
  DA(D.req, S); DA(D.n_moves, S); DA(D.path_len, S); DA(D.path, S * PATH_STRIDE);
  DA(D.bucket, S); DA(D.urow, S); DA(D.nn_row, S); DA(D.row_slot, S);
  DA(D.root_Qp, S); DA(D.root_Qn, S);
  DA(D.c_sims, S); DA(D.c_evals, S); DA(D.c_term, S); DA(D.c_depth, S);
  DA(D.blocks, S * 2 * (size_t)D.cap);
  DA(D.table, T);
  uint64_t *gid, *p0, *p1;
  DA(gid, R); DA(p0, R); DA(p1, R);
  D.game_id = gid; D.p0 = p0; D.p1 = p1;
  DA(D.n_samples, R); DA(D.s_mask, R * MAXS); DA(D.s_value, R * MAXS);
  DA(D.s_policy, R * MAXS * 7); DA(D.s_qp, R * MAXS); DA(D.s_qn, R * MAXS);
  DA(D.g, 1); DA(D.movers, S);
  DA(e->scratch4, 4); DA(e->row_mask, S); DA(e->row_value, S); DA(e->row_model, S);
  cudaError_t err = cudaMemset(D.state, 0, S * sizeof(uint32_t));
  if (err == cudaSuccess) err = cudaMemset(D.g, 0, sizeof(Globals));
  if (err == cudaSuccess) err = cudaMemset(D.table, 0, T * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMallocHost((void**)&e->h_globals, sizeof(Globals));
  if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->h_status, sizeof(HostStatus), cudaHostAllocMapped);
  if (err == cudaSuccess) {
    memset((void*)e->h_status, 0, sizeof(HostStatus));
    err = cudaHostGetDevicePointer((void**)&D.status, (void*)e->h_status, 0);
  }
  if (err != cudaSuccess) {
    c4a0_engine_destroy(e);
    return fail(C4A0_E_CUDA, "engine init failed: %s", cudaGetErrorString(err));
  }
  *out = e;
  return 0;
}

void c4a0_engine_destroy(c4a0_engine* e) {
  if (!e) return;
  cudaSetDevice(e->cfg.device);
  cudaDeviceSynchronize();
  for (void* p : e->allocs) cudaFree(p);
  if (e->h_globals) cudaFreeHost(e->h_globals);
  if (e->h_status) cudaFreeHost((void*)e->h_status);
  for (auto ev : e->ev)
    if (ev) cudaEventDestroy(ev);
  delete e;
}

size_t c4a0_engine_device_bytes(const c4a0_engine* e) { return e ? e->bytes : 0; }

int c4a0_engine_bind_io(c4a0_engine* e, void* planes, const float* logits, const float* qp,
                        const float* qn) {
  if (!e || !planes || !logits || !qp || !qn) return fail(C4A0_E_INVALID, "null argument");
  if (((uintptr_t)planes & 15u) != 0) return fail(C4A0_E_INVALID, "planes_dev must be 16-byte aligned");
  e->D.planes = planes;
  e->D.logits = logits;
  e->D.qp = qp;
  e->D.qn = qn;
  e->b_logits = const_cast<float*>(logits);
  e->b_qp = const_cast<float*>(qp);
  e->b_qn = const_cast<float*>(qn);
  e->io_bound = true;
  return 0;
}

int c4a0_engine_set_requests(c4a0_engine* e, const uint64_t* game_id, const uint64_t* p0,
                             const uint64_t* p1, uint32_t n, void* stream) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (!e->io_bound) return fail(C4A0_E_INVALID, "bind_io() must precede set_requests()");
  if (n > e->cfg.max_requests) return fail(C4A0_E_INVALID, "%u requests exceed max_requests=%u", n, e->cfg.max_requests);
  if (n && (!game_id || !p0 || !p1)) return fail(C4A0_E_INVALID, "null request arrays");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaSetDevice(e->cfg.device));
  Dev& D = e->D;
  if (n) {
    CK(cudaMemcpyAsync((void*)D.game_id, game_id, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync((void*)D.p0, p0, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemcpyAsync((void*)D.p1, p1, n * sizeof(uint64_t), cudaMemcpyHostToDevice, s));
    CK(cudaMemsetAsync(D.n_samples, 0, n * sizeof(uint32_t), s));
    // unused sample cells read back as zeros
    CK(cudaMemsetAsync(D.s_mask, 0, (size_t)n * MAXS * 8, s));
    CK(cudaMemsetAsync(D.s_value, 0, (size_t)n * MAXS * 8, s));
    CK(cudaMemsetAsync(D.s_policy, 0, (size_t)n * MAXS * 28, s));
    CK(cudaMemsetAsync(D.s_qp, 0, (size_t)n * MAXS * 4, s));
    CK(cudaMemsetAsync(D.s_qn, 0, (size_t)n * MAXS * 4, s));
  }
  CK(cudaMemsetAsync(D.table, 0, ((size_t)D.table_mask + 1) * sizeof(unsigned long long), s));
  k_init<<<blocks_for(D.n_slots, 256), 256, 0, s>>>(D, n);
  k_scan<<<1, SCAN_THREADS, 0, s>>>(D);
  k_pack<<<blocks_for((size_t)D.n_slots * 8, 256), 256, 0, s>>>(D);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(s));  // the host arrays may be freed by the caller after return
  e->n_req = n;
  e->have_requests = true;
  e->steps = 0;
  return 0;
}

int c4a0_engine_step(c4a0_engine* e, void* stream) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (!e->have_requests) return fail(C4A0_E_INVALID, "set_requests() must precede step()");
  return launch_tick(e, (cudaStream_t)stream, nullptr);
}

int c4a0_engine_step_timed(c4a0_engine* e, void* stream, float* ms_step, float* ms_move) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (!e->have_requests) return fail(C4A0_E_INVALID, "set_requests() must precede step()");
  cudaStream_t s = (cudaStream_t)stream;
  if (!e->ev[0])
    for (int i = 0; i < 3; i++) CK(cudaEventCreate(&e->ev[i]));
  int r = launch_tick(e, s, e->ev);
  if (r) return r;
  CK(cudaStreamSynchronize(s));
  float a = 0, b = 0;
  CK(cudaEventElapsedTime(&a, e->ev[0], e->ev[1]));
  CK(cudaEventElapsedTime(&b, e->ev[1], e->ev[2]));
  if (ms_step) *ms_step = a;
  if (ms_move) *ms_move = b;
  return 0;
}

int c4a0_engine_eval_builtin(c4a0_engine* e, int kind, void* stream) {
  if (!e || !e->io_bound) return fail(C4A0_E_INVALID, "engine not bound");
  if (kind != C4A0_EVAL_UNIFORM && kind != C4A0_EVAL_HASH) return fail(C4A0_E_INVALID, "bad evaluator kind");
  k_eval_builtin<<<blocks_for(e->D.n_slots, 256), 256, 0, (cudaStream_t)stream>>>(e->D, kind, e->b_logits, e->b_qp, e->b_qn);
  CK(cudaGetLastError());
  return 0;
}

int c4a0_engine_poll(c4a0_engine* e, c4a0_progress* out, void* stream) {
  if (!e || !out) return fail(C4A0_E_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(e->h_globals, e->D.g, sizeof(Globals), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const Globals& g = *e->h_globals;
  out->n_requests = g.n_req;
  out->n_started = g.next_req < g.n_req ? g.next_req : g.n_req;
  out->n_finished = g.n_finished;
  out->n_running = g.n_running;
  out->n_movers = g.n_movers;
  out->n_rows = g.n_rows;
  out->error = g.error;
  if (g.error) return fail(C4A0_E_ENGINE, "%s", kEngineErr);
  return 0;
}

int c4a0_engine_stats(c4a0_engine* e, c4a0_stats* out, void* stream) {
  if (!e || !out) return fail(C4A0_E_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemsetAsync(e->scratch4, 0, 4 * sizeof(unsigned long long), s));
  k_sum_counters<<<64, 256, 0, s>>>(e->D, e->scratch4);
  CK(cudaGetLastError());
  unsigned long long h4[4];
  CK(cudaMemcpyAsync(h4, e->scratch4, sizeof(h4), cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(e->h_globals, e->D.g, sizeof(Globals), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  const Globals& g = *e->h_globals;
  out->sims = h4[0];
  out->nn_evals = g.rows_total;
  out->leaf_requests = g.leaves_total;
  out->terminal_leaf_sims = h4[2];
  out->select_depth_sum = h4[3];
  out->expansions = h4[1];
  out->skipped_root_sims = g.skipped_root_sims;
  out->moves = g.moves;
  out->samples = g.samples;
  out->steps = e->steps;
  out->compacted_blocks = g.compacted_blocks;
  return 0;
}

int c4a0_engine_fetch_rows(c4a0_engine* e, uint32_t* n_rows, uint64_t* leaf_mask, uint64_t* leaf_value,
                           uint64_t* model_id, void* stream) {
  if (!e || !n_rows) return fail(C4A0_E_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  size_t S = e->D.n_slots;
  k_gather_rows<<<blocks_for(S, 256), 256, 0, s>>>(e->D, e->row_mask, e->row_value, e->row_model);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(e->h_globals, e->D.g, sizeof(Globals), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  uint32_t n = e->h_globals->n_rows;
  *n_rows = n;
  if (n && leaf_mask) CK(cudaMemcpyAsync(leaf_mask, e->row_mask, n * 8, cudaMemcpyDeviceToHost, s));
  if (n && leaf_value) CK(cudaMemcpyAsync(leaf_value, e->row_value, n * 8, cudaMemcpyDeviceToHost, s));
  if (n && model_id) CK(cudaMemcpyAsync(model_id, e->row_model, n * 8, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int c4a0_engine_fetch_results(c4a0_engine* e, uint32_t first, uint32_t n, uint32_t* n_samples,
                              uint64_t* mask, uint64_t* value, float* policy, float* qp, float* qn,
                              void* stream) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if ((uint64_t)first + n > e->n_req) return fail(C4A0_E_INVALID, "result range out of bounds");
  cudaStream_t s = (cudaStream_t)stream;
  const Dev& D = e->D;
  size_t o = (size_t)first * MAXS, c = (size_t)n * MAXS;
  if (n_samples) CK(cudaMemcpyAsync(n_samples, D.n_samples + first, n * 4, cudaMemcpyDeviceToHost, s));
  if (mask) CK(cudaMemcpyAsync(mask, D.s_mask + o, c * 8, cudaMemcpyDeviceToHost, s));
  if (value) CK(cudaMemcpyAsync(value, D.s_value + o, c * 8, cudaMemcpyDeviceToHost, s));
  if (policy) CK(cudaMemcpyAsync(policy, D.s_policy + o * 7, c * 7 * 4, cudaMemcpyDeviceToHost, s));
  if (qp) CK(cudaMemcpyAsync(qp, D.s_qp + o, c * 4, cudaMemcpyDeviceToHost, s));
  if (qn) CK(cudaMemcpyAsync(qn, D.s_qn + o, c * 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  return 0;
}

int c4a0_engine_results_dev(c4a0_engine* e, uint32_t** n_samples, uint64_t** mask, uint64_t** value,
                            float** policy, float** qp, float** qn) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (n_samples) *n_samples = e->D.n_samples;
  if (mask) *mask = e->D.s_mask;
  if (value) *value = e->D.s_value;
  if (policy) *policy = e->D.s_policy;
  if (qp) *qp = e->D.s_qp;
  if (qn) *qn = e->D.s_qn;
  return 0;
}

int c4a0_engine_slot_info(c4a0_engine* e, uint32_t slot, c4a0_slot_info* out, void* stream) {
  if (!e || !out) return fail(C4A0_E_INVALID, "null argument");
  if (slot >= e->D.n_slots) return fail(C4A0_E_INVALID, "slot out of range");
  cudaStream_t s = (cudaStream_t)stream;
  const Dev& D = e->D;
  uint32_t na;
  CK(cudaMemcpyAsync(&out->state, D.state + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->request, D.req + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->n_moves, D.n_moves + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->root_visits, D.root_N + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->root_mask, D.root_mask + slot, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->root_value, D.root_value + slot, 8, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->root_q_sum_penalty, D.root_Qp + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->root_q_sum_no_penalty, D.root_Qn + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&na, D.n_alloc + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&out->nn_row, D.nn_row + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out->n_blocks = na ? na - 1 : 0;
  return 0;
}

namespace {
struct Dumper {
  const Block* blocks;
  uint32_t* buf;
  size_t cap, w;
  void put(uint32_t v) {
    if (w < cap) buf[w] = v;
    w++;
  }
  void children(uint32_t b, Pos pos) {
    unsigned legal = c4::legal_mask(pos.mask);
    const Block& B = blocks[b];
    for (int c = 0; c < 7; c++) {
      if (!((legal >> c) & 1u)) {
        for (int k = 0; k < 5; k++) put(0);
        continue;
      }
      put(B.child[c] ? 2u : 1u);
      put(B.N[c]);
      put(c4::f32_bits(B.Qp[c]));
      put(c4::f32_bits(B.Qn[c]));
      put(c4::f32_bits(B.P[c]));
      if (B.child[c]) children(B.child[c], c4::make_move(pos, c));
    }
  }
};
}  // namespace

int c4a0_engine_dump_tree(c4a0_engine* e, uint32_t slot, uint32_t* buf, size_t cap, size_t* needed,
                          void* stream) {
  if (!e || !needed) return fail(C4A0_E_INVALID, "null argument");
  c4a0_slot_info info;
  int r = c4a0_engine_slot_info(e, slot, &info, stream);
  if (r) return r;
  cudaStream_t s = (cudaStream_t)stream;
  const Dev& D = e->D;
  uint32_t half, rb;
  CK(cudaMemcpyAsync(&half, D.half + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaMemcpyAsync(&rb, D.root_block + slot, 4, cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  std::vector<Block> host(info.n_blocks + 1);
  const Block* src = D.blocks + ((size_t)slot * 2 + half) * D.cap;
  CK(cudaMemcpyAsync(host.data(), src, host.size() * sizeof(Block), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  Dumper d{host.data(), buf, buf ? cap : 0, 0};
  d.put(rb ? 2u : 1u);
  d.put(info.root_visits);
  d.put(c4::f32_bits(info.root_q_sum_penalty));
  d.put(c4::f32_bits(info.root_q_sum_no_penalty));
  if (rb) d.children(rb, Pos{info.root_mask, info.root_value});
  *needed = d.w;
  return 0;
}

// ------------------------------------------------------------------------------------------------
// The host loop of self_play() (self_play.rs:60-129 spawns threads and waits for done_queue; here
// the host only sequences two kinds of GPU work per engine):
//     tree tick (our kernels)  ->  read n_rows from mapped host memory  ->  network graph for the
//     smallest bucket >= n_rows (a cudaGraphExec_t the caller captured, e.g. with torch)
// With two engines (two half-batches on two streams) one engine's tree tick and host round trip
// hide under the other engine's network.
// ------------------------------------------------------------------------------------------------
int c4a0_engine_run(c4a0_engine* const* engines, uint32_t n_engines, const c4a0_nn_graph* const* graphs,
                    const uint32_t* n_graphs, void* const* streams, uint64_t max_ticks,
                    uint32_t time_kernels_every, c4a0_run_report* out) {
  if (!engines || !graphs || !n_graphs || !streams || !out || n_engines == 0 || n_engines > 8)
    return fail(C4A0_E_INVALID, "bad argument");
  memset(out, 0, sizeof(*out));
  struct Lane {
    c4a0_engine* e;
    cudaStream_t s;
    uint32_t expect;   // status tick the host waits for next
    bool done;
    cudaEvent_t t0, t1;
  };
  std::vector<Lane> lanes(n_engines);
  std::vector<cudaEvent_t> kev;  // triples
  for (uint32_t i = 0; i < n_engines; i++) {
    c4a0_engine* e = engines[i];
    if (!e || !e->have_requests) return fail(C4A0_E_INVALID, "engine %u has no requests", i);
    if (n_graphs[i] == 0 || !graphs[i]) return fail(C4A0_E_INVALID, "engine %u has no network graphs", i);
    for (uint32_t k = 0; k < n_graphs[i]; k++) {
      if (!graphs[i][k].graph_exec) return fail(C4A0_E_INVALID, "null graph_exec");
      if (k && graphs[i][k].rows <= graphs[i][k - 1].rows) return fail(C4A0_E_INVALID, "network graphs must be sorted by rows");
    }
    if (graphs[i][n_graphs[i] - 1].rows < e->D.n_slots)
      return fail(C4A0_E_INVALID, "the largest network graph must cover n_slots rows");
    lanes[i] = Lane{e, (cudaStream_t)streams[i], e->h_status->tick, e->n_req == 0, nullptr, nullptr};
    CK(cudaEventCreate(&lanes[i].t0));
    CK(cudaEventCreate(&lanes[i].t1));
  }
  auto cleanup = [&]() {
    for (auto& L : lanes) {
      if (L.t0) cudaEventDestroy(L.t0);
      if (L.t1) cudaEventDestroy(L.t1);
    }
    for (auto ev : kev) cudaEventDestroy(ev);
  };
  auto wall0 = std::chrono::steady_clock::now();
  int rc = 0;
  // prime: the planes of the initial roots are already packed (set_requests) -> first network call
  auto launch_nn = [&](uint32_t i) -> int {
    Lane& L = lanes[i];
    uint32_t rows = L.e->h_status->n_rows;
    const c4a0_nn_graph* g = graphs[i];
    uint32_t k = 0;
    while (k + 1 < n_graphs[i] && g[k].rows < rows) k++;
    CK(cudaGraphLaunch((cudaGraphExec_t)g[k].graph_exec, L.s));
    out->nn_launches++;
    out->bucket_launches[k < 31 ? k : 31]++;
    out->nn_rows_launched += g[k].rows;
    return 0;
  };
  auto launch_tree = [&](uint32_t i) -> int {
    Lane& L = lanes[i];
    cudaEvent_t* ev = nullptr;
    if (time_kernels_every && (L.e->steps % time_kernels_every) == 0 && kev.size() < 3 * 4096) {
      size_t b = kev.size();
      for (int q = 0; q < 3; q++) {
        cudaEvent_t x;
        CK(cudaEventCreate(&x));
        kev.push_back(x);
      }
      ev = &kev[b];
    }
    int r = launch_tick(L.e, L.s, ev);
    if (r) return r;
    L.expect++;
    out->ticks++;
    return 0;
  };
  for (uint32_t i = 0; i < n_engines && !rc; i++) {
    Lane& L = lanes[i];
    CK(cudaEventRecord(L.t0, L.s));
    if (L.done) continue;
    rc = launch_nn(i);
    if (!rc) rc = launch_tree(i);
  }
  uint32_t remaining = 0;
  for (auto& L : lanes) remaining += L.done ? 0 : 1;
  uint32_t cur = 0;
  while (remaining && !rc) {
    Lane& L = lanes[cur];
    if (!L.done) {
      // wait for the tick's status (k_scan wrote it through mapped memory)
      uint64_t spins = 0;
      while (L.e->h_status->tick != L.expect) {
        if (++spins > 2000) {
          if (cudaStreamQuery(L.s) == cudaSuccess && L.e->h_status->tick != L.expect) {
            rc = fail(C4A0_E_CUDA, "tick status never arrived (stream idle)");
            break;
          }
          std::this_thread::yield();
        }
      }
      if (rc) break;
      std::atomic_thread_fence(std::memory_order_acquire);
      if (L.e->h_status->error) {
        rc = fail(C4A0_E_ENGINE, "%s", kEngineErr);
        break;
      }
      if (L.e->h_status->n_finished >= L.e->n_req) {
        L.done = true;
        remaining--;
        CK(cudaEventRecord(L.t1, L.s));
      } else if (max_ticks && out->ticks >= max_ticks) {
        rc = fail(C4A0_E_INVALID, "max_ticks reached before all games finished");
        break;
      } else {
        rc = launch_nn(cur);
        if (!rc) rc = launch_tree(cur);
      }
    }
    cur = (cur + 1) % n_engines;
  }
  for (auto& L : lanes) {
    if (rc) break;
    if (L.e->n_req == 0) CK(cudaEventRecord(L.t1, L.s));
  }
  for (auto& L : lanes) cudaStreamSynchronize(L.s);
  if (!rc) {
    float mx = 0;
    for (auto& L : lanes) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, L.t0, L.t1) == cudaSuccess && ms > mx) mx = ms;
    }
    out->device_ms = mx;
    double a = 0, b = 0;
    uint32_t n = 0;
    for (size_t q = 0; q + 2 < kev.size(); q += 3) {
      float x = 0, y = 0;
      if (cudaEventElapsedTime(&x, kev[q], kev[q + 1]) == cudaSuccess &&
          cudaEventElapsedTime(&y, kev[q + 1], kev[q + 2]) == cudaSuccess) {
        a += x;
        b += y;
        n++;
      }
    }
    out->kernel_samples = n;
    out->k_step_ms_sum = a;
    out->k_move_ms_sum = b;
  }
  out->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  cleanup();
  return rc;
}

}  // extern "C"
