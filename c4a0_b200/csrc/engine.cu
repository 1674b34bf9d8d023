// engine.cu — lockstep MCTS self-play on sm_100a: kernels + the engine half of the C-ABI
// (include/c4a0_engine.h).
//
// What the reference does per game with a heap tree of Rc<RefCell<Node>> (rust/src/mcts.rs:332-355)
// and a thread pool (rust/src/self_play.rs), this file does for thousands of games at once with
// block-structured trees in HBM:
//
//   * A tree is a bump-allocated array of 160-byte BLOCKS.  A block belongs to one expanded node
//     and holds its 7 children: one 16-byte record {N, Qp, Qn, P} per child plus child[8] block
//     indices, i.e. exactly the fields UCT selection reads for one level (mcts.rs:359-388) — one
//     vector load per lane — and a backup updates a child with one 16-byte read-modify-write.
//     Node positions are never stored: selection replays the moves on bitboards.
//   * Eight lanes own one game for a whole tick (k_step; four games per warp, warp-uniform control
//     flow): they consume the network's answer (mask + softmax + expand + backup, mcts.rs:83-155),
//     play the move when the root reached n_iterations (temperature + seeded sample + re-root,
//     self_play.rs:283-300, mcts.rs:187-222), finish / re-seat the game, and select the next leaf
//     (mcts.rs:160-183; lane c evaluates child c, argmax = 3 shuffle steps + a ballot with the
//     reference's last-maximum tie-break).  A tick is a dependent chain per game; the 7-way UCT
//     math, the softmax and the move sampling are the parts that shorten when spread over lanes.
//   * Backup walks the path recorded by selection: one f32 add per node per simulation in
//     simulation order — the reference's accumulation order.
//   * Re-rooting keeps the chosen child's subtree and drops the siblings.  Nothing is copied while
//     the arena has room (root_block simply becomes the child's block); when a half fills up, one
//     CTA copies the live subtree breadth-first into the other half: a CTA of k_step that is done
//     with its own games takes the arena off an epoch-tagged list while slower CTAs still run
//     (work stealing), long backlogs go to k_tail.  Memory per game is 2 * arena_blocks * 160 B,
//     with arena_blocks >= n_iterations + 2.
//   * The network batch is dense and de-duplicated, like the reference's NN thread builds it
//     (self_play.rs:203-220: HashSet<Pos> per model).  As soon as a game knows its next leaf it
//     inserts (position, model) into an epoch-tagged hash table; the first game to claim a key
//     draws the next row number from an atomic counter and writes the input planes of that row,
//     later games with the same key just remember who leads it.  Rows are dense, 0..n_rows-1.
//   * Optionally (C4A0_FLAG_EVAL_CACHE) every answer of the network is kept, for the duration of a
//     job, in a 2-way set-associative table keyed by (position, model): a leaf the job has evaluated before
//     is answered inside the tick and the game goes on to its next simulation.  On top of that
//     (C4A0_FLAG_SPECULATE) small batches are topped up with the children of the leaves being
//     expanded, whose answers land in the same table before selection gets to them.  Both assume an
//     evaluator that is a pure function of (model, position); game records do not change.
//
//   * Optionally (c4a0_config.dirichlet_*) root priors get Dirichlet noise; the reference has none, and
//     the default kernel k_step<false> contains none of that code.
//
// One tick = k_step, then the network on rows [0, n_rows).  The last CTA of k_step to finish closes
// the tick: it compacts the (few) arenas still on the list and publishes the tick's status (n_rows,
// finished games) to mapped host memory.  Only a long backlog of compactions is handed to a second
// kernel, k_tail (one CTA per arena), which the host launches when the status word asks for it.
// The host loop (c4a0_engine_run / c4a0_engine_run_net at the end of this file) keeps one tick queued
// ahead of the GPU; with the library's own network kernel (net.cu) the two kernels of a tick are chained
// by programmatic dependent launch.
// No CPU fallback exists: every entry point that computes needs the GPU and fails loudly.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>  // header-only; ranges cost a pointer check unless a profiler is attached
#include <stddef.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <chrono>
#include <thread>
#include <vector>

#include "c4_math.cuh"
#include "c4_rng.cuh"
#include "c4_rules.cuh"
#include "../../include/c4a0_net.h"
#include "common.cuh"

namespace {

using c4::Pos;
using c4host::blocks_for;
using c4host::fail;

// ------------------------------------------------------------------------------------------------
// Data layout
// ------------------------------------------------------------------------------------------------
struct __align__(16) ChildStat {  // one 16-byte vector per child: a backup touches one sector
  uint32_t N;  // visit_count            (mcts.rs:335)
  float Qp;    // q_sum_penalty          (mcts.rs:336)
  float Qn;    // q_sum_no_penalty       (mcts.rs:337)
  float P;     // initial_policy_value   (mcts.rs:338)
};
struct __align__(32) Block {
  ChildStat rec[8];   // child c of the node (entry 7 is padding)
  uint32_t child[8];  // block index of child c's own block, 0 = child not expanded
};
static_assert(sizeof(Block) == 160, "block must be ten 16-byte vectors");

// Evaluation cache (C4A0_FLAG_EVAL_CACHE): one network answer per 64-byte entry; the two entries of a 128-byte
// line form a set (2-way set associative, see cache_lookup).
//   tag = job << 33 | dying << 32 | tick.  `job` counts set_requests() calls, so entries of earlier
//   jobs are simply stale.  An entry is READABLE in tick e iff job matches, dying == 0 and tick < e.
//   State changes go through one atomicCAS on the tag and never touch a payload that a reader of
//   the same tick could accept:
//     claim (stale, or dying since an earlier tick)  -> {job, 0, e + 1}; the claimer writes the key
//          now and the payload when its answer arrives in tick e + 1; readable from tick e + 2.
//          During tick e + 1 itself a reader that finds `row` set takes the answer straight from
//          the network's output row (those buffers do not change within a tick).
//     kill  (readable, other key)                    -> {job, 1, e}; payload untouched; claimable
//          from tick e + 1 (always-replace with one tick of delay)
struct __align__(64) EvalEntry {
  uint64_t key;            // pos_key(): 49 bits that identify the position
  unsigned long long tag;
  uint64_t model;
  float qp, qn;
  float logit[7];
  uint32_t row;            // 1 + the network row that evaluates the position in the batch of the claiming tick
                           // (0 = not known to the claimer): the answer can be read from the network's
                           // output buffers one tick before the payload is readable
};
static_assert(sizeof(EvalEntry) == 64, "an entry is two sectors");

// Everything a game needs besides its tree, one 128-byte line per slot.
struct __align__(128) Slot {
  uint64_t root_mask, root_value;
  uint64_t leaf_mask, leaf_value;  // the leaf waiting for the network (= its de-duplication key ...
  uint64_t model0, model1;         //   ... with the model to move: player0_id at even ply, mcts.rs:70-76)
  uint32_t root_N;
  float root_Qp, root_Qn;
  uint32_t root_block;             // block of the root's children, 0 = root not expanded
  uint32_t n_alloc;                // next free block index of the live arena half (index 0 unused)
  uint32_t path_len;
  uint32_t state;
  uint32_t half;
  uint32_t req;                    // request index of the game seated here
  uint32_t n_moves;
  uint32_t nn_leader;              // slot that leads this game's leaf key: its rowtag names the network row
  uint32_t cache_own;              // 1 + the evaluation-cache entry this game claimed for its waiting leaf, 0 = none
  unsigned long long c_sims, c_exp, c_term, c_depth;
};
__host__ __device__ inline uint64_t leaf_model_of(const Slot& S) {
  return (c4::popc64(S.leaf_mask) & 1) ? S.model1 : S.model0;
}
static_assert(sizeof(Slot) == 128, "slot state must be one cache line");
// load_game()/store_game() move the line as 16-byte vectors: V[0] root, V[1] leaf, V[2] models,
// V[3] root statistics, V[4] allocator/state, V[5] request/moves/leader, V[6..7] counters
static_assert(offsetof(Slot, leaf_mask) == 16 && offsetof(Slot, model0) == 32 && offsetof(Slot, root_N) == 48 &&
                  offsetof(Slot, n_alloc) == 64 && offsetof(Slot, req) == 80 && offsetof(Slot, c_sims) == 96,
              "Slot layout changed: update the vector views");

constexpr int PATH_STRIDE = 48;  // <= 42 levels below a root; six entries per lane
constexpr int MAXS = C4A0_MAX_SAMPLES;
constexpr int STEP_THREADS = 128;
enum : uint32_t { ST_IDLE = 0, ST_WAIT_NN = 1, ST_CONTINUE = 2, ST_NEED_MOVE = 3 };

struct Globals {  // one instance in device memory
  uint32_t tick;  // epoch of the hash table / scan: advanced at the end of every k_tail
  uint32_t n_req;
  uint32_t next_req;
  uint32_t n_finished;
  uint32_t n_running;
  uint32_t n_movers;
  uint32_t n_rows;      // rows of the batch the network is about to evaluate
  uint32_t rows_acc;    // row counter of the tick being built
  uint32_t wait_acc;    // games waiting for the network in the tick being built
  uint32_t tail_done;   // k_tail CTA ticket
  uint32_t step_done;   // k_step CTA ticket
  uint32_t tail_pending;  // k_step left compaction work (and the closing of the tick) to k_tail
  int32_t error;
  uint32_t mover_claim;   // entries of `movers` already taken by a CTA of k_step (compaction work stealing)
  unsigned long long skipped_root_sims;
  unsigned long long moves;
  unsigned long long samples;
  unsigned long long compacted_blocks;
  unsigned long long compactions;
  unsigned long long rows_total;    // sum over ticks of n_rows (positions actually evaluated)
  unsigned long long leaves_total;  // sum over ticks of games waiting for the network
  unsigned long long cache_hits;    // leaves answered by the evaluation cache
  unsigned long long cache_inserts; // entries claimed for a network answer
  unsigned long long spec_total;    // rows evaluated speculatively (children of expanded leaves)
  // speculation (C4A0_FLAG_SPECULATE): spare rows of a small batch evaluate children of expanded leaves
  uint32_t spec_budget;    // speculative rows this tick may draw (set when the previous tick closed)
  uint32_t spec_acc;       // ... drawn so far (may overshoot the budget; the excess is dropped)
  uint32_t spec_ok;        // ... that became rows of the batch being built
  uint32_t spec_count[2];  // entries of spec_list[parity] written in the tick of that parity
};

struct HostStatus {  // mapped pinned host memory, written by k_tail at the end of every tick
  volatile uint32_t tick;
  volatile uint32_t n_rows;
  volatile uint32_t n_finished;
  volatile uint32_t n_running;
  volatile uint32_t n_movers;
  volatile int32_t error;
  volatile uint32_t need_tail;  // == epoch: k_step finished but many arenas need compaction, launch k_tail
};

struct Dev {  // passed to kernels by value
  uint32_t n_slots, n_iter, cap, max_inline, plane_bf16, plane_stride, dedup, table_mask;
  float c_expl, c_pen;
  float dir_alpha, dir_eps;  // Dirichlet noise on root priors; dir_eps == 0: off (the reference's behaviour)
  Slot* slots;
  uint32_t* path;   // [n_slots][PATH_STRIDE]: (block << 3 | column) per level of the selected path
  Block* blocks;    // [n_slots][2][cap]
  uint32_t* row_slot;   // [row_cap] slot whose leaf is the row; 0xffffffff for a speculative row
  uint64_t* row_model;  // [row_cap] model that has to evaluate the row (mcts.rs:70-76)
  unsigned long long* rowtag;  // [2][n_slots] epoch << 32 | row, written by the slot that leads a key; the half is
                               // the epoch's parity: a leader may publish its next leaf while followers of its last
                               // one (warps that start later in the same tick) still look its row up
  // per request
  const uint64_t *game_id, *p0, *p1;
  uint32_t* n_samples;
  uint64_t *s_mask, *s_value;
  float *s_policy, *s_qp, *s_qn;
  // global
  Globals* g;
  HostStatus* status;  // device address of the mapped host struct
  unsigned long long* movers;      // [n_slots] epoch << 32 | slot: games whose arena half is full, in arrival order
  unsigned long long* table;       // [table_mask+1]: epoch << 32 | leader slot
  EvalEntry* cache;                // [cache_mask+1] evaluation cache, nullptr = off
  uint32_t cache_mask, job;
  const float* ln_tab;             // [ln_n] c4_logf((float)i): visit counts are small integers
  uint32_t ln_n;
  uint32_t spec_cap;               // rows a batch may be topped up to with speculative evaluations, 0 = off
  uint32_t spec_thr;               // ... while the rows games ask for are at most this many
  uint32_t row_cap;                // rows of the I/O buffers: n_slots (+ spec_cap with speculation)
  uint32_t max_inline_spec;        // in-kernel simulation budget while speculation is running
  uint2* spec_list;                // [2][spec_cap] (row, cache entry) of the speculative rows of a batch
  uint64_t *row_mask, *row_value;  // [row_cap] position of every row of the batch being built
  // NN io
  void* planes;
  const float *logits, *qp, *qn;
  // optional per-slot phase timing of k_step (8 x u32 cycles per slot), see c4a0_engine_debug_phases
  uint32_t* dbg;
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

// NN input planes of `p` into row `row` (c4r.rs:378-392): 84 values as 16-byte vectors.
__device__ __forceinline__ void write_planes(const Dev& D, uint32_t row, Pos p) {
  const uint64_t mine = p.mask & p.value, theirs = p.mask & ~p.value;
  // bit i of `bits` (i < 84) is plane element i
  const uint64_t lo = mine | (theirs << 42);  // elements 0..63
  const uint64_t hi = theirs >> 22;           // elements 64..83
  if (D.plane_bf16) {
    __nv_bfloat16* base = reinterpret_cast<__nv_bfloat16*>(D.planes) + (size_t)row * D.plane_stride;
    if ((D.plane_stride & 7u) == 0u) {  // rows are 16-byte aligned: 8 bf16 per store
      uint4* dst = reinterpret_cast<uint4*>(base);
#pragma unroll
      for (int v = 0; v < 11; v++) {  // elements 84..87 of the last vector are padding (zeros)
        uint32_t b = (uint32_t)((v < 8 ? lo >> (8 * v) : hi >> (8 * (v - 8))) & 0xffull);
        if (v == 10) b &= 0x0fu;
        uint4 w;
        w.x = ((b & 1u) ? 0x3f80u : 0u) | ((b & 2u) ? 0x3f800000u : 0u);  // bf16(1.0) = 0x3f80
        w.y = ((b & 4u) ? 0x3f80u : 0u) | ((b & 8u) ? 0x3f800000u : 0u);
        w.z = ((b & 16u) ? 0x3f80u : 0u) | ((b & 32u) ? 0x3f800000u : 0u);
        w.w = ((b & 64u) ? 0x3f80u : 0u) | ((b & 128u) ? 0x3f800000u : 0u);
        dst[v] = w;
      }
    } else {  // e.g. the natural stride 84: rows are 8-byte aligned, 4 bf16 per store
      uint2* dst = reinterpret_cast<uint2*>(base);
#pragma unroll
      for (int v = 0; v < 21; v++) {
        uint32_t b = (uint32_t)((v < 16 ? lo >> (4 * v) : hi >> (4 * (v - 16))) & 0xfull);
        dst[v] = make_uint2(((b & 1u) ? 0x3f80u : 0u) | ((b & 2u) ? 0x3f800000u : 0u),
                            ((b & 4u) ? 0x3f80u : 0u) | ((b & 8u) ? 0x3f800000u : 0u));
      }
    }
  } else {
    float4* dst = reinterpret_cast<float4*>(reinterpret_cast<float*>(D.planes) + (size_t)row * D.plane_stride);
#pragma unroll
    for (int v = 0; v < 21; v++) {
      uint32_t b = (uint32_t)((v < 16 ? lo >> (4 * v) : hi >> (4 * (v - 16))) & 0xfull);
      dst[v] = make_float4((float)(b & 1u), (float)((b >> 1) & 1u), (float)((b >> 2) & 1u), (float)((b >> 3) & 1u));
    }
  }
}

// Leaf de-duplication + row assignment (self_play.rs:203-220).  Called by ONE lane per game once
// its slot line (leaf position + model ids) has been stored.  Table entry = epoch << 32 | leader
// slot; an entry of another epoch is empty, so the table is never cleared.  The first game to claim
// a key leads it: it takes the next row of the batch and writes the planes; the others only record
// the leader.  (Which of several equal leaves leads, and hence the order of rows, depends on
// arrival order; rows are independent in the network, so results do not.)
__device__ __forceinline__ void publish_leaf(const Dev& D, uint32_t slot, uint64_t km, uint64_t kv, uint64_t kmod,
                                             uint32_t epoch, uint32_t cache_own) {
  uint32_t leader = slot;
  if (D.dedup) {
    __threadfence();  // the slot line must be visible before the slot can be found in the table
    uint32_t h = (uint32_t)splitmix64(km * 0x9E3779B97F4A7C15ULL ^ splitmix64(kv ^ kmod)) & D.table_mask;
    const unsigned long long mine = ((unsigned long long)epoch << 32) | slot;
    for (;;) {
      unsigned long long* e = D.table + h;
      unsigned long long cur = *reinterpret_cast<volatile unsigned long long*>(e);
      if ((uint32_t)(cur >> 32) != epoch) {
        unsigned long long prev = atomicCAS(e, cur, mine);
        if (prev == cur) break;  // first game with this key in this tick: we lead
        cur = prev;
        if ((uint32_t)(cur >> 32) != epoch) continue;
      }
      const Slot* Ld = D.slots + (uint32_t)cur;
      const uint64_t lm = __ldcg(&Ld->leaf_mask), lv = __ldcg(&Ld->leaf_value);
      const uint64_t lmod = (c4::popc64(lm) & 1) ? __ldcg(&Ld->model1) : __ldcg(&Ld->model0);
      if (lm == km && lv == kv && lmod == kmod) {
        leader = (uint32_t)cur;
        break;
      }
      h = (h + 1) & D.table_mask;  // another key lives here: linear probing
    }
  }
  D.slots[slot].nn_leader = leader;
  if (leader == slot) {
    const uint32_t row = atomicAdd(&D.g->rows_acc, 1u);
    D.rowtag[(size_t)(epoch & 1u) * D.n_slots + slot] = ((unsigned long long)epoch << 32) | row;
    D.row_slot[row] = slot;
    D.row_model[row] = kmod;
    D.row_mask[row] = km;
    D.row_value[row] = kv;
    if (cache_own) D.cache[cache_own - 1u].row = row + 1u;
    write_planes(D, row, Pos{km, kv});
  }
}

// ------------------------------------------------------------------------------------------------
// Eight lanes per game, four games per warp.  Lane c (< 7) owns child / column c of whatever node
// the game is looking at; lane 7 rides along.  ALL control flow is warp-uniform: every lane of the
// warp executes every shuffle (constant full mask, width 8), and per-game decisions are predicates.
// What matters for a tick is the length of the dependent chain per game, and the 7-way UCT math,
// the softmax and the move sampling are exactly the parts that parallelise across the lanes.
// ------------------------------------------------------------------------------------------------
constexpr unsigned FULL = 0xffffffffu;

struct Lanes {
  int l;           // 0..7 within the game
  unsigned gbase;  // first lane of the game within the warp (0, 8, 16, 24)
};
__device__ __forceinline__ Lanes make_lanes() {
  int lane = threadIdx.x & 31;
  return Lanes{lane & 7, (unsigned)(lane & 24)};
}
template <typename T>
__device__ __forceinline__ T gshfl(T v, int src) {
  return __shfl_sync(FULL, v, src, 8);
}
template <typename T>
__device__ __forceinline__ T gxor(T v, int m) {
  return __shfl_xor_sync(FULL, v, m, 8);
}
__device__ __forceinline__ unsigned gballot(const Lanes& L, bool p) {
  return (__ballot_sync(FULL, p) >> L.gbase) & 0xffu;
}
// ((((((v0+v1)+v2)+v3)+v4)+v5)+v6): the reference's iter().sum() over the seven columns
__device__ __forceinline__ float fold7(float v) {
  float s = gshfl(v, 0);
#pragma unroll
  for (int i = 1; i < 7; i++) s = s + gshfl(v, i);
  return s;
}
__device__ __forceinline__ float gmax8(float v) {
#pragma unroll
  for (int m = 1; m < 8; m <<= 1) v = fmaxf(v, gxor(v, m));
  return v;
}

// Per-game working set: group-uniform values, held redundantly by the 8 lanes.
struct Game {
  uint32_t slot, state;
  Pos root, leaf;
  uint32_t rootN, root_block, n_alloc, len, half, req, n_moves, nn_leader;
  float rootQp, rootQn;
  Block* arena;
  uint32_t* path;
  uint32_t pr[6];  // this lane's share of the selected path: entries l, l+8, ..., l+40
  uint32_t sims, exps, term, depth;  // this tick's contribution to the slot's counters
  uint32_t cache_own, hits, claims;  // evaluation cache: entry owed a payload (+1); this tick's hits / claims
  bool reseated;                     // a new game took the slot during this tick (model ids change)
};

__device__ __forceinline__ uint32_t pr_get(const uint32_t (&pr)[6], uint32_t k) {
  uint32_t v = pr[0];
#pragma unroll
  for (int i = 1; i < 6; i++) v = (k == (uint32_t)i) ? pr[i] : v;
  return v;
}
__device__ __forceinline__ void pr_set(uint32_t (&pr)[6], uint32_t k, uint32_t v) {
#pragma unroll
  for (int i = 0; i < 6; i++) pr[i] = (k == (uint32_t)i) ? v : pr[i];
}

// One dependent memory round trip: the slot line (same address for the 8 lanes = one broadcast
// transaction per 16-byte vector) and this lane's path entries are fetched together.  Only what a
// tick needs in registers is loaded: the model ids and the counter totals stay in memory.
__device__ __forceinline__ void load_game(const Dev& D, const Lanes& L, uint32_t slot, Game& G) {
  const uint4* V = reinterpret_cast<const uint4*>(D.slots + slot);
  G.slot = slot;
  G.path = D.path + (size_t)slot * PATH_STRIDE;
#pragma unroll
  for (int k = 0; k < 6; k++) G.pr[k] = G.path[L.l + 8 * k];
  const uint4 v0 = V[0], v1 = V[1], v3 = V[3], v4 = V[4], v5 = V[5];
  G.root.mask = (uint64_t)v0.x | ((uint64_t)v0.y << 32);
  G.root.value = (uint64_t)v0.z | ((uint64_t)v0.w << 32);
  G.leaf.mask = (uint64_t)v1.x | ((uint64_t)v1.y << 32);
  G.leaf.value = (uint64_t)v1.z | ((uint64_t)v1.w << 32);
  G.rootN = v3.x;
  G.rootQp = __uint_as_float(v3.y);
  G.rootQn = __uint_as_float(v3.z);
  G.root_block = v3.w;
  G.n_alloc = v4.x;
  G.len = v4.y;
  G.state = v4.z;
  G.half = v4.w;
  G.req = v5.x;
  G.n_moves = v5.y;
  G.nn_leader = v5.z;
  G.cache_own = v5.w;
  G.arena = arena_of(D, slot, G.half);
  G.sims = G.exps = G.term = G.depth = 0;
  G.hits = G.claims = 0;
  G.reseated = false;
}
// One lane writes the line back (five 16-byte vectors) and adds the tick's counters to the totals.
__device__ __forceinline__ void store_game(const Dev& D, const Game& G, uint32_t state) {
  Slot* S = D.slots + G.slot;
  uint4* V = reinterpret_cast<uint4*>(S);
  V[0] = make_uint4((uint32_t)G.root.mask, (uint32_t)(G.root.mask >> 32), (uint32_t)G.root.value, (uint32_t)(G.root.value >> 32));
  V[1] = make_uint4((uint32_t)G.leaf.mask, (uint32_t)(G.leaf.mask >> 32), (uint32_t)G.leaf.value, (uint32_t)(G.leaf.value >> 32));
  if (G.reseated) {
    S->model0 = D.p0[G.req];
    S->model1 = D.p1[G.req];
  }
  V[3] = make_uint4(G.rootN, __float_as_uint(G.rootQp), __float_as_uint(G.rootQn), G.root_block);
  V[4] = make_uint4(G.n_alloc, G.len, state, G.half);
  V[5] = make_uint4(G.req, G.n_moves, G.nn_leader, G.cache_own);
  if (G.sims) {
    ulonglong2* C = reinterpret_cast<ulonglong2*>(&S->c_sims);
    ulonglong2 a = C[0], b = C[1];
    a.x += G.sims;
    a.y += G.exps;
    b.x += G.term;
    b.y += G.depth;
    C[0] = a;
    C[1] = b;
  }
}
// the model that has to evaluate the slot's stored leaf (mcts.rs:70-76); read after store_game()
__device__ __forceinline__ uint64_t stored_leaf_model(const Dev& D, const Game& G) {
  const Slot* S = D.slots + G.slot;
  return (c4::popc64(G.leaf.mask) & 1) ? S->model1 : S->model0;
}

// mcts.rs:137-155 — add (qp, qn) at the leaf, alternate the sign towards the root.  The path nodes
// are distinct, so lanes update them in parallel: one f32 add per node per simulation.
__device__ __forceinline__ void backup(const Lanes& L, Game& G, bool pred, float qp, float qn) {
  __syncwarp();  // earlier block writes of the other lanes are visible
  if (pred) {
#pragma unroll
    for (int k = 0; k < 6; k++) {
      const uint32_t j = L.l + 8 * k;
      if (j < G.len) {
        const uint32_t e = G.pr[k];
        Block* B = G.arena + (e >> 3);
        const uint32_t c = e & 7u;
        const bool neg = ((G.len - 1 - j) & 1u) != 0;
        uint4* rp = reinterpret_cast<uint4*>(&B->rec[c]);  // one 16-byte read-modify-write
        uint4 r = *rp;
        r.x += 1u;
        r.y = __float_as_uint(__uint_as_float(r.y) + (neg ? -qp : qp));
        r.z = __float_as_uint(__uint_as_float(r.z) + (neg ? -qn : qn));
        *rp = r;
      }
    }
    const bool neg = (G.len & 1u) != 0;
    G.rootN += 1u;
    G.rootQp += neg ? -qp : qp;
    G.rootQn += neg ? -qn : qn;
  }
  __syncwarp();  // statistics visible to the selection that follows
}

// mcts.rs:359-388: -(Qp/(N+1)) + c * (sqrt(ln(N_parent)/(N+1)) * (P + 1e-8)), f32, no contraction
// The zero tests do not change any result (0/x = 0 keeps its sign, sqrt(+0) = +0): they keep the very
// common zero numerators (unvisited children: Qp = 0; parents with one visit: ln 1 = 0) off the slow
// path of the IEEE division / square root sequences.
__device__ __forceinline__ float div_exact(float a, float b) { return a == 0.0f ? a : a / b; }  // b > 0
__device__ __forceinline__ float uct(uint32_t n, float qs, float pr, float lnp, float c_expl) {
  float nf = (float)n + 1.0f;
  float q = div_exact(qs, nf);
  float ex = lnp == 0.0f ? 0.0f : sqrtf(lnp / nf);
  ex = ex * (pr + 1e-8f);
  return (-q) + (c_expl * ex);
}

// order-preserving map f32 -> u32 (never 0), with -0.0 folded onto +0.0 like a float comparison
__device__ __forceinline__ uint32_t ordkey(float u) {
  uint32_t b = __float_as_uint(u + 0.0f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}

// mcts.rs:160-183 — walk from the root to a node without children; records the path.  One tree
// level per iteration for every game of the warp that is still descending.
__device__ __forceinline__ Pos select_leaf(const Dev& D, const Lanes& L, Game& G, bool pred) {
  Pos pos = G.root;
  uint32_t np = G.rootN, b = pred ? G.root_block : 0u, len = 0;
  while (__any_sync(FULL, b != 0u)) {
    const bool act = b != 0u;
    uint32_t n = 0u, ch = 0u;
    float qs = 0.0f, pr = 0.0f;
    if (act) {
      const Block* B = G.arena + b;
      const uint4 r = *reinterpret_cast<const uint4*>(&B->rec[L.l]);
      n = r.x;
      qs = __uint_as_float(r.y);
      pr = __uint_as_float(r.w);
      ch = B->child[L.l];
    }
    const unsigned legal = c4::legal_mask(pos.mask);
    // ln(parent visits), mcts.rs:378-379; the table holds the same function's values
    const float lnp = np < D.ln_n ? __ldg(D.ln_tab + np) : c4::c4_logf((float)np);
    const float u = uct(n, qs, pr, lnp, D.c_expl);
    const bool ok = act && L.l < 7 && ((legal >> L.l) & 1u);
    const uint32_t key = ok ? ordkey(u) : 0u;
    uint32_t m = key;
#pragma unroll
    for (int k = 1; k < 8; k <<= 1) m = max(m, gxor(m, k));
    const unsigned winners = gballot(L, ok && key == m);
    const int best = 31 - __clz(winners);  // max_by_key: the LAST maximum wins; -1 if none
    const int src = best & 7;
    const uint32_t bn = gshfl(n, src), bc = gshfl(ch, src);
    if (act) {
      if (best >= 0 && len < 42u) {
        if ((len & 7u) == (uint32_t)L.l) {  // the lane that owns this level keeps and persists it
          const uint32_t e = (b << 3) | (uint32_t)best;
          pr_set(G.pr, len >> 3, e);
          G.path[len] = e;
        }
        len++;
        pos = c4::make_move(pos, best);
        np = bn;
        b = bc;
      } else {
        b = 0u;  // cannot happen: an expanded node has a legal move and a tree is <= 42 plies deep
      }
    }
  }
  if (pred) G.len = len;
  return pos;
}

// ---- Dirichlet noise on the root's priors (c4a0_config.dirichlet_alpha / _epsilon) -----------------------
// Not in the reference (mcts.rs:114-132 keeps the masked softmax as it is): off unless configured.  eta ~ Dir(alpha)
// over the legal moves = normalised Gamma(alpha, 1) draws; lane c draws for column c from a counter-based stream
// keyed by (game_id, moves played, column), so the noise of a game does not depend on scheduling, slots or batches.
// Gamma: Marsaglia-Tsang squeeze (alpha < 1: Gamma(alpha + 1) * U^(1/alpha)), normals by the polar method; the
// f32 arithmetic is the kernel's usual exact kind (no contraction, c4_logf / c4_expf), so it is reproducible.
struct NoiseRng {
  uint64_t s;
  __device__ __forceinline__ uint32_t next() {
    s += 0x9E3779B97F4A7C15ULL;
    uint64_t z = s;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return (uint32_t)((z ^ (z >> 31)) >> 32);
  }
  __device__ __forceinline__ float uniform() { return ((float)(next() >> 8) + 0.5f) * (1.0f / 16777216.0f); }  // in (0, 1)
};
__device__ __noinline__ float gamma_draw(float alpha, uint64_t seed) {
  NoiseRng r{seed};
  const float a = alpha < 1.0f ? alpha + 1.0f : alpha;
  const float d = a - 1.0f / 3.0f, c = 1.0f / sqrtf(9.0f * d);
  float g = d;  // (the mean, should 64 proposals all be rejected: probability < 1e-30)
  for (int it = 0; it < 64; it++) {
    const float u = 2.0f * r.uniform() - 1.0f, v = 2.0f * r.uniform() - 1.0f;
    const float s2 = u * u + v * v;
    if (s2 >= 1.0f || s2 == 0.0f) continue;
    const float x = u * sqrtf(-2.0f * c4::c4_logf(s2) / s2);
    const float t = 1.0f + c * x;
    if (t <= 0.0f) continue;
    const float v3 = t * t * t;
    if (c4::c4_logf(r.uniform()) < 0.5f * x * x + d - d * v3 + d * c4::c4_logf(v3)) {
      g = d * v3;
      break;
    }
  }
  if (alpha < 1.0f) g = g * c4::c4_expf(c4::c4_logf(r.uniform()) / alpha);
  return g;
}
// Called by the whole warp; `pred` marks the games whose root priors are being set, `p` is this lane's prior.
// (scalars only: a reference to the kernel's parameter struct would force a local copy of it)
__device__ __noinline__ float root_noise(float alpha, float eps, int l, bool pred, unsigned legal, uint64_t game_id,
                                         uint32_t n_moves, float p) {
  const bool act = pred && l < 7 && ((legal >> l) & 1u);
  float g = 0.0f;
  if (act) g = gamma_draw(alpha, splitmix64(game_id ^ 0xD1B54A32D192ED03ULL) ^ (((uint64_t)n_moves << 8 | (uint64_t)l) * 0xA24BAED4963EE407ULL));
  __syncwarp();
  const float sum = fold7(g);
  return (act && sum > 0.0f) ? (1.0f - eps) * p + eps * (g / sum) : p;
}

// mask_policy + softmax + expand_leaf + backup for a leaf whose evaluation (x = this lane's logit,
// vq, vn) has arrived from the network or from the evaluation cache (c4r.rs:272-286,
// mcts.rs:416-434, 114-132, 137-155).  Returns false for a game whose arena overflowed (engine bug;
// reported).
template <bool NOISE>
__device__ __forceinline__ bool apply_answer(const Dev& D, const Lanes& L, Game& G, bool pred, float xin, float vq,
                                             float vn) {
  const unsigned legal = c4::legal_mask(G.leaf.mask);
  const bool ok = pred && L.l < 7 && ((legal >> L.l) & 1u);
  const float x = ok ? xin : -c4::f32_inf();
  const float mx = gmax8(x);
  const float e = ok ? c4::c4_expf(x - mx) : 0.0f;
  const float s = fold7(e);
  float p = ok ? div_exact(e, s) : 0.0f;
  if (NOISE && __any_sync(FULL, pred && G.len == 0u))  // this leaf is the root of its game's search
    p = root_noise(D.dir_alpha, D.dir_eps, L.l, pred && G.len == 0u, legal, (pred && G.len == 0u) ? D.game_id[G.req] : 0ull, G.n_moves, p);
  const uint32_t nb = G.n_alloc;
  const bool fits = nb < D.cap;
  if (pred && fits) {
    G.n_alloc = nb + 1u;
    Block* B = G.arena + nb;  // expand_leaf: one new block holding the 7 children
    *reinterpret_cast<uint4*>(&B->rec[L.l]) = make_uint4(0u, 0u, 0u, __float_as_uint(p));
    B->child[L.l] = 0u;
    if (G.len == 0) G.root_block = nb;
    G.depth += G.len;
    G.sims++;
    G.exps++;
  }
  // the new block hangs under the last node of the path, which lane (len-1)&7 remembers
  const uint32_t last = G.len ? G.len - 1u : 0u;
  const uint32_t e2 = gshfl(pr_get(G.pr, last >> 3), (int)(last & 7u));
  if (pred && fits && G.len != 0u && L.l == 0) G.arena[e2 >> 3].child[e2 & 7u] = nb;
  backup(L, G, pred && fits, pred ? vq : 0.0f, pred ? vn : 0.0f);
  return !pred || fits;
}

// ---- evaluation cache ---------------------------------------------------------------------------
using c4::pos_key;  // c4_rules.cuh: the 49-bit key of the evaluation cache
constexpr unsigned long long TAG_DYING = 1ull << 32;

// The game's freshly selected, non-terminal leaf: answered from the cache (-> true, values in
// x/vq/vn) or not.  On a miss lane 0 may claim the entry for the answer the network will deliver
// next tick (G.cache_own), or mark a readable entry of another key for replacement.
// The table is 2-way set associative: the two 64-byte entries of a 128-byte line form a set (one memory
// transaction serves both).  Measured on the bench job with one way: in the last 500 ticks of a job EVERY miss
// of a live game was a collision (the wanted position's entry held another live key), and each collision costs
// two network round trips (mark dying, ask uncached, claim on the next request).
struct WayState {
  unsigned long long tag;
  bool mine, dying, old, same;
};
__device__ __forceinline__ WayState way_state(const uint4 a, const uint64_t emodel, uint64_t key, uint64_t model, uint32_t job,
                                              uint32_t epoch) {
  WayState s;
  const uint64_t ekey = (uint64_t)a.x | ((uint64_t)a.y << 32);
  s.tag = (unsigned long long)a.z | ((unsigned long long)a.w << 32);
  s.mine = (uint32_t)(s.tag >> 33) == job;
  s.dying = (s.tag & TAG_DYING) != 0ull;
  s.old = (uint32_t)s.tag < epoch;  // the last state change happened in an earlier tick
  s.same = ekey == key && emodel == model;
  return s;
}
// way to claim for a key that has no entry in the set: an empty / stale way first, else one whose entry has been
// dying since an earlier tick; -1 = both ways hold live entries
__device__ __forceinline__ int claimable_way(const WayState& s0, const WayState& s1) {
  if (!s0.mine) return 0;
  if (!s1.mine) return 1;
  if (s0.dying && s0.old) return 0;
  if (s1.dying && s1.old) return 1;
  return -1;
}

__device__ __forceinline__ bool cache_lookup(const Dev& D, const Lanes& L, Game& G, bool pred, Pos leaf, uint64_t model,
                                             uint32_t epoch, float& x, float& vq, float& vn) {
  const uint64_t key = pos_key(leaf);
  const uint32_t h0 = (uint32_t)splitmix64(key ^ (model * 0x9E3779B97F4A7C15ULL)) & D.cache_mask & ~1u;
  bool hit = false;
  if (pred) {
    EvalEntry* E = D.cache + h0;  // ways E[0], E[1]
    const int li = L.l < 7 ? L.l : 6;
    const uint4 a0 = *reinterpret_cast<const uint4*>(E), b0 = *(reinterpret_cast<const uint4*>(E) + 1);          // key, tag | model, qp, qn
    const uint4 a1 = *reinterpret_cast<const uint4*>(E + 1), b1 = *(reinterpret_cast<const uint4*>(E + 1) + 1);
    const float lg0 = E[0].logit[li], lg1 = E[1].logit[li];
    const WayState s0 = way_state(a0, (uint64_t)b0.x | ((uint64_t)b0.y << 32), key, model, D.job, epoch);
    const WayState s1 = way_state(a1, (uint64_t)b1.x | ((uint64_t)b1.y << 32), key, model, D.job, epoch);
    const bool hit0 = s0.mine && !s0.dying && s0.old && s0.same, hit1 = s1.mine && !s1.dying && s1.old && s1.same;
    // claimed during the last tick with a known row: that row of the last batch holds the answer
    uint32_t frow = 0u;
    if (!hit0 && !hit1) {
      if (s0.mine && !s0.dying && s0.same && (uint32_t)s0.tag == epoch) frow = E[0].row;
      else if (s1.mine && !s1.dying && s1.same && (uint32_t)s1.tag == epoch) frow = E[1].row;
    }
    if (hit0 || hit1) {
      hit = true;
      x = hit0 ? lg0 : lg1;
      vq = __uint_as_float(hit0 ? b0.z : b1.z);
      vn = __uint_as_float(hit0 ? b0.w : b1.w);
      G.hits++;
    } else if (frow) {
      hit = true;
      x = D.logits[(size_t)(frow - 1u) * 7 + li];
      vq = D.qp[frow - 1u];
      vn = D.qn[frow - 1u];
      G.hits++;
    } else if (L.l == 0) {
      const unsigned long long jb = (unsigned long long)D.job << 33;
      const bool here0 = s0.mine && s0.same, here1 = s1.mine && s1.same;  // the key has an entry, not readable (yet)
      int w = -1;
      if (here0 || here1) {
        const WayState& s = here0 ? s0 : s1;
        if (s.dying && s.old) w = here0 ? 0 : 1;  // its own dying copy may be claimed again
      } else {
        w = claimable_way(s0, s1);
      }
      if (w >= 0) {
        EvalEntry* Ew = E + w;
        const unsigned long long tag = w ? s1.tag : s0.tag;
        if (atomicCAS(&Ew->tag, tag, jb | (unsigned long long)(epoch + 1u)) == tag) {
          Ew->key = key;
          Ew->model = model;
          Ew->row = 0u;  // set by publish_leaf() if this game also leads the key
          G.cache_own = h0 + (uint32_t)w + 1u;
          G.claims++;
        }
      } else if (!here0 && !here1) {
        // both ways hold other live keys: retire the readable one that changed longest ago (claimable next tick)
        const bool r0 = !s0.dying && s0.old, r1 = !s1.dying && s1.old;
        int k = -1;
        if (r0 && r1) k = (uint32_t)s0.tag <= (uint32_t)s1.tag ? 0 : 1;
        else if (r0) k = 0;
        else if (r1) k = 1;
        if (k >= 0) atomicCAS(&E[k].tag, k ? s1.tag : s0.tag, jb | TAG_DYING | (unsigned long long)epoch);
      }
    }
  }
  return hit;
}

// the model that has to evaluate `leaf` for the game in G (mcts.rs:70-76)
__device__ __forceinline__ uint64_t model_to_play(const Dev& D, const Game& G, Pos leaf) {
  const bool odd = (c4::popc64(leaf.mask) & 1) != 0;
  if (G.reseated) return odd ? D.p1[G.req] : D.p0[G.req];
  const Slot* S = D.slots + G.slot;
  return odd ? S->model1 : S->model0;
}

// Speculation: the leaf G.leaf has just been expanded.  While the batch being built is small, its
// spare rows evaluate the children of such leaves ahead of time: lane c looks child c up and, if the
// position is neither cached nor already on its way, claims its cache entry exactly like a game that
// misses (tag tick = e + 1) and draws a row of the batch for it.  Nobody waits for these rows; the
// answers are moved into the claimed entries during the next tick (spec_collect) and are readable
// from tick e + 2, when selection may arrive at the child and finds it answered.
__device__ __forceinline__ bool spec_request(const Dev& D, const Game& G, Pos pos, uint32_t epoch, uint32_t budget) {
  const uint64_t model = model_to_play(D, G, pos);
  const uint64_t key = pos_key(pos);
  const uint32_t h0 = (uint32_t)splitmix64(key ^ (model * 0x9E3779B97F4A7C15ULL)) & D.cache_mask & ~1u;
  EvalEntry* E = D.cache + h0;
  // (L2 loads: other SMs claim entries during the tick, and a stale L1 line would only cost a failed CAS)
  const uint4 a0 = __ldcg(reinterpret_cast<const uint4*>(E)), a1 = __ldcg(reinterpret_cast<const uint4*>(E + 1));
  const WayState s0 = way_state(a0, __ldcg(reinterpret_cast<const unsigned long long*>(&E[0].model)), key, model, D.job, epoch);
  const WayState s1 = way_state(a1, __ldcg(reinterpret_cast<const unsigned long long*>(&E[1].model)), key, model, D.job, epoch);
  // cached or on its way already
  if ((s0.mine && s0.same && !(s0.dying && s0.old)) || (s1.mine && s1.same && !(s1.dying && s1.old))) return true;
  const int w = claimable_way(s0, s1);
  if (w < 0) return true;  // both ways hold live entries: a guess never evicts
  Globals* g = D.g;
  const uint32_t s = atomicAdd(&g->spec_acc, 1u);
  if (s >= budget) return false;
  uint2 item = make_uint2(0xffffffffu, 0u);
  EvalEntry* Ew = E + w;
  const unsigned long long tag = w ? s1.tag : s0.tag;
  if (atomicCAS(&Ew->tag, tag, ((unsigned long long)D.job << 33) | (unsigned long long)(epoch + 1u)) == tag) {
    Ew->key = key;
    Ew->model = model;
    const uint32_t row = atomicAdd(&g->rows_acc, 1u);
    atomicAdd(&g->spec_ok, 1u);
    Ew->row = row + 1u;
    D.row_slot[row] = 0xffffffffu;
    D.row_model[row] = model;
    D.row_mask[row] = pos.mask;
    D.row_value[row] = pos.value;
    write_planes(D, row, pos);
    item = make_uint2(row, h0 + (uint32_t)w);
  }
  D.spec_list[(size_t)(epoch & 1u) * D.spec_cap + s] = item;
  return true;
}
__device__ __forceinline__ void speculate_children(const Dev& D, const Lanes& L, const Game& G, bool pred, uint32_t epoch,
                                                   uint32_t budget) {
  const unsigned legal = c4::legal_mask(G.leaf.mask);
  const int c = L.l < 7 ? L.l : 0;
  const Pos child = c4::make_move(G.leaf, c);
  float tq0, tq1;
  // (also asking for the position a game visits first below each child — its last legal successor, since a
  // node with one visit scores all children equal — was measured and does not shorten the job)
  if (pred && L.l < 7 && ((legal >> c) & 1u) && c4::terminal_value(child, D.c_pen, &tq0, &tq1) == c4::NONE)
    spec_request(D, G, child, epoch, budget);
}

// The answers of last tick's speculative rows go into the cache entries claimed for them (any time
// during this tick: the entries become readable in the next one).  Grid-stride over the list.
__device__ __forceinline__ void spec_collect(const Dev& D, uint32_t epoch) {
  const uint32_t par = (epoch - 1u) & 1u;
  const uint32_t n = D.g->spec_count[par];
  for (uint32_t k = blockIdx.x * blockDim.x + threadIdx.x; k < n; k += gridDim.x * blockDim.x) {
    const uint2 it = D.spec_list[(size_t)par * D.spec_cap + k];
    if (it.x == 0xffffffffu) continue;
    EvalEntry* E = D.cache + it.y;
#pragma unroll
    for (int i = 0; i < 7; i++) E->logit[i] = D.logits[(size_t)it.x * 7 + i];
    E->qp = D.qp[it.x];
    E->qn = D.qn[it.x];
  }
}

__device__ __forceinline__ void seat_game(const Dev& D, Game& G, uint32_t r) {
  G.reseated = true;
  G.root = Pos{0ull, 0ull};
  G.rootN = 0u;
  G.rootQp = 0.0f;
  G.rootQn = 0.0f;
  G.root_block = 0u;
  G.n_alloc = 1u;
  G.req = r;
  G.n_moves = 0u;
  G.len = 0u;
}

enum MoveResult : int { MV_CONTINUE = 0, MV_IDLE = 1, MV_COMPACT = 2 };

// The root reached n_iterations (self_play.rs:283-313): sample and play a move (mcts.rs:187-222) or,
// when that ends the game, emit its samples (mcts.rs:271-313) and seat the next request.  Executed
// by the whole warp; `pred` marks the games that actually move.
template <bool NOISE>
__device__ __forceinline__ int play_move(const Dev& D, const Lanes& L, Game& G, bool pred) {
  Globals* g = D.g;
  const int l = L.l;
  const float uniform = 1.0f / 7.0f;
  const float ninf = -c4::f32_inf();
  const uint32_t rb = pred ? G.root_block : 0u;
  uint32_t n_c = 0u, ch = 0u;
  float q_p = 0.0f, q_n = 0.0f;
  if (rb) {
    const Block* RB = G.arena + rb;
    const uint4 r = *reinterpret_cast<const uint4*>(&RB->rec[l]);
    n_c = r.x;
    q_p = __uint_as_float(r.y);
    q_n = __uint_as_float(r.z);
    ch = RB->child[l];
  }
  // root_policy (mcts.rs:396-412): visit counts of the children, normalised
  const float cnt = (float)n_c;
  const float sum = fold7(cnt);
  const float pol = (rb == 0u || sum == 0.0f) ? uniform : div_exact(cnt, sum);
  // apply_temperature (mcts.rs:439-454); T from self_play.rs:294-299
  const float T = c4::temperature_for_ply(c4::ply(G.root.mask));
  const float p0 = gshfl(pol, 0);
  const bool alleq = gballot(L, l == 7 || pol == p0) == 0xffu;
  const float lg = c4::c4_logf(pol) / T;
  const float ex = (l < 7) ? c4::c4_expf(lg) : 0.0f;
  const float lse = c4::c4_logf(fold7(ex));
  float v = c4::c4_expf(lg - lse);
  v = v < 0.0f ? 0.0f : v;
  v = v > 1.0f ? 1.0f : v;
  const float pmx = gmax8(l < 7 ? pol : ninf);
  const float one = (l < 7 && pol == pmx) ? 1.0f : 0.0f;
  const float v0 = div_exact(one, fold7(one));
  float w = (T == 1.0f || alleq) ? pol : (T == 0.0f ? v0 : v);
  if (l == 7) w = 0.0f;
  // WeightedIndex::new(w) (cumulative left fold) and Uniform<f32>[0, total)
  float run = 0.0f, cum = 0.0f;
#pragma unroll
  for (int i = 0; i < 7; i++) {
    float wi = gshfl(w, i);
    run = (i == 0) ? wi : run + wi;
    if (i == l) cum = run;
  }
  const float total = run;
  const bool bad = gballot(L, l < 7 && !(w >= 0.0f)) != 0u || !(total > 0.0f) || total == c4::f32_inf();
  float scale = total;
  const float max_rand = c4::bits_f32(0x3f7ffffeu);
  if (!bad)
    while (scale * max_rand + 0.0f >= total) scale = c4::bits_f32(c4::f32_bits(scale) - 1u);
  uint32_t key[8];
  const uint64_t gid = pred ? D.game_id[G.req] : 0ull;
  c4::seed_to_key(c4::move_seed(gid, (int)G.n_moves), key);
  const uint32_t u32 = c4::chacha12_first_word(key);
  const float x = (c4::bits_f32((u32 >> 9) | 0x3f800000u) - 1.0f) * scale + 0.0f;
  const int col = __popc(gballot(L, l < 6 && cum <= x));  // partition_point(|c| c <= x) over 6 entries
  const unsigned legal = c4::legal_mask(G.root.mask);
  const bool err = pred && (bad || col > 6 || !((legal >> col) & 1u) || rb == 0u || G.n_moves >= 42u);
  const bool go = pred && !err;
  // the chosen child: position, statistics, subtree
  const int src = col & 7;
  const uint32_t newN = gshfl(n_c, src), child = gshfl(ch, src);
  const float newQp = gshfl(q_p, src), newQn = gshfl(q_n, src);
  const Pos np = c4::make_move(G.root, col > 6 ? 0 : col);
  float tqp, tqn;
  const int t = c4::terminal_value(np, D.c_pen, &tqp, &tqn);
  const bool fin = go && t != c4::NONE;
  // the next request, should the game end here (self_play.rs:55-58 queues them all up front)
  uint32_t r = 0u;
  if (fin && l == 0) r = atomicAdd(&g->next_req, 1u);
  r = gshfl(r, 0);
  int result = MV_CONTINUE;
  if (err) {
    if (l == 0) {
      g->error = C4A0_E_ENGINE;  // the reference panics here (mcts.rs:190-197): park the game, report
      atomicSub(&g->n_running, 1u);
    }
    result = MV_IDLE;
  }
  if (go) {
    // RecordedMove (mcts.rs:198-203) goes straight into the sample store
    const size_t s0 = (size_t)G.req * MAXS;
    size_t si = s0 + G.n_moves;
    if (l < 7) D.s_policy[si * 7 + l] = pol;
    if (l == 0) {
      D.s_mask[si] = G.root.mask;
      D.s_value[si] = G.root.value;
      atomicAdd(&g->moves, 1ull);
    }
    G.n_moves += 1u;
    if (fin) {
      // to_result (mcts.rs:271-313): alternate the terminal value back through the moves
      const uint32_t Lm = G.n_moves;
      for (uint32_t k = l; k < Lm; k += 8) {
        bool neg = ((Lm - k) & 1u) != 0;
        D.s_qp[s0 + k] = neg ? -tqp : tqp;
        D.s_qn[s0 + k] = neg ? -tqn : tqn;
      }
      si = s0 + Lm;
      if (l < 7) D.s_policy[si * 7 + l] = uniform;
      if (l == 0) {
        D.s_mask[si] = np.mask;
        D.s_value[si] = np.value;
        D.s_qp[si] = tqp;
        D.s_qn[si] = tqn;
        D.n_samples[G.req] = Lm + 1u;
        atomicAdd(&g->samples, (unsigned long long)(Lm + 1u));
        atomicAdd(&g->n_finished, 1u);
        // the reference keeps simulating the terminal root until N >= n
        if (newN < D.n_iter) atomicAdd(&g->skipped_root_sims, (unsigned long long)(D.n_iter - newN));
      }
      if (r < g->n_req) {
        seat_game(D, G, r);
      } else {
        if (l == 0) atomicSub(&g->n_running, 1u);
        result = MV_IDLE;
      }
    } else {
      // re-root (mcts.rs:200-205): the child's subtree and statistics are kept
      G.root = np;
      G.rootN = newN;
      G.rootQp = newQp;
      G.rootQn = newQn;
      G.root_block = child;
      G.len = 0u;
      // until the next move at most n_iter - N expansions happen: do they still fit in this half?
      const uint32_t need = (D.n_iter > newN ? D.n_iter - newN : 0u) + 1u;
      if (G.n_alloc + need > D.cap) result = MV_COMPACT;
    }
  }
  // a reused subtree becomes the root of the next search: its priors get their noise now (see root_noise)
  const bool renoise = go && !fin && child != 0u;
  if (NOISE && __any_sync(FULL, renoise)) {
    ChildStat* R = &G.arena[renoise ? child : 0u].rec[l];
    const float p0n = renoise ? R->P : 0.0f;
    const float p1n = root_noise(D.dir_alpha, D.dir_eps, l, renoise, c4::legal_mask(np.mask), gid, G.n_moves, p0n);
    if (renoise && l < 7) R->P = p1n;
  }
  return result;
}

// Advance the (up to four) games of this warp until each needs the network (WAIT_NN), has used its
// in-kernel budget of simulations that need no network row (CONTINUE), needs its tree compacted
// (NEED_MOVE) or holds no game (IDLE).  `running` marks the lanes of live games, `pend` those that
// hold an evaluation (x, vq, vn) of their leaf G.leaf that still has to be applied: the network's
// answer on entry, an evaluation-cache hit later on.  Returns the new state.
template <bool NOISE>
__device__ __forceinline__ uint32_t run_games(const Dev& D, const Lanes& L, Game& G, bool running, uint32_t state,
                                              uint32_t epoch, uint32_t spec_budget, uint32_t max_inline, bool pend, float x,
                                              float vq, float vn) {
  uint32_t inl = 0;
  for (;;) {
    if (__any_sync(FULL, pend)) {
      const bool ok = apply_answer<NOISE>(D, L, G, pend, x, vq, vn);
      if (!ok) {
        if (L.l == 0) D.g->error = C4A0_E_ENGINE;
        running = false;
      }
      if (spec_budget) speculate_children(D, L, G, pend && ok, epoch, spec_budget);
      pend = false;
    }
    const bool need_move = running && G.rootN >= D.n_iter;  // self_play.rs:283: after every simulation
    if (__any_sync(FULL, need_move)) {
      const int r = play_move<NOISE>(D, L, G, need_move);
      if (need_move && r == MV_IDLE) {
        running = false;
        state = ST_IDLE;
      }
      if (need_move && r == MV_COMPACT) {
        running = false;
        state = ST_NEED_MOVE;
      }
      continue;
    }
    if (running && inl >= max_inline) {
      running = false;
      state = ST_CONTINUE;
    }
    if (!__any_sync(FULL, running)) break;
    const Pos leaf = select_leaf(D, L, G, running);
    float tqp, tqn;
    const int t = c4::terminal_value(leaf, D.c_pen, &tqp, &tqn);
    const bool term = running && t != c4::NONE;
    const bool ask = running && t == c4::NONE;
    if (ask) G.leaf = leaf;
    if (D.cache != nullptr && __any_sync(FULL, ask)) {
      // a position this job has evaluated before: the stored answer is applied in this tick
      pend = cache_lookup(D, L, G, ask, leaf, ask ? model_to_play(D, G, leaf) : 0ull, epoch, x, vq, vn);
      if (pend) inl++;
    }
    if (ask && !pend) {
      running = false;
      state = ST_WAIT_NN;
    }
    if (__any_sync(FULL, term)) {
      // terminal leaf: mcts.rs:92-98 — no expansion, back up the objective value
      if (term) {
        G.depth += G.len;
        G.sims++;
        G.term++;
        inl++;
      }
      backup(L, G, term, tqp, tqn);
    }
  }
  return state;
}

// Called by lane 0 of a game whose slot line is stored and whose warp has finished with the tree (the caller
// put a __syncwarp() behind the other lanes' writes): from here on any CTA may compact the arena.  The entry
// carries the epoch, so the list never needs clearing and a reader can tell a written entry from a stale one.
__device__ __forceinline__ void push_mover(const Dev& D, uint32_t slot, uint32_t epoch) {
  __threadfence();
  const uint32_t i = atomicAdd(&D.g->n_movers, 1u);
  atomicExch(D.movers + i, ((unsigned long long)epoch << 32) | slot);
}
constexpr uint32_t NO_MOVER = 0xffffffffu;
__device__ __forceinline__ uint32_t mover_slot(const Dev& D, uint32_t m, uint32_t epoch) {
  // the entry is written right after the counter was advanced: a handful of polls at most.  Bounded all the
  // same — a wait that long is a protocol bug, and a bug must surface as C4A0_E_ENGINE, never as a hung GPU
  for (uint32_t spin = 0; spin < (1u << 22); spin++) {
    const unsigned long long v = *reinterpret_cast<volatile unsigned long long*>(D.movers + m);
    if ((uint32_t)(v >> 32) == epoch) {
      __threadfence();
      return (uint32_t)v;
    }
  }
  D.g->error = C4A0_E_ENGINE;
  return NO_MOVER;
}

constexpr uint32_t STEAL_BACKLOG_MAX = 16;   // k_step's CTAs take compactions off the list while at most this many wait
constexpr uint32_t INLINE_COMPACTIONS = 4;  // up to this many, k_step's last CTA compacts by itself

// ------------------------------------------------------------------------------------------------
// Compaction: the CTA copies the live subtree of one game breadth-first into the other half of its
// arena (level by level; a level's blocks are contiguous in the destination).  The game resumes in
// the next tick (state CONTINUE).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void compact_one(const Dev& D, uint32_t slot) {
  __shared__ uint32_t sh_next, sh_head;
  if (slot >= D.n_slots) return;  // NO_MOVER (CTA-uniform): see mover_slot
  Slot* S = D.slots + slot;
  const uint32_t half = S->half;
  const uint32_t src_root = S->root_block;
  Block* src = arena_of(D, slot, half);
  Block* dst = arena_of(D, slot, half ^ 1u);
  __syncthreads();
  if (threadIdx.x == 0) {
    sh_head = 1u;
    sh_next = src_root ? 2u : 1u;
  }
  if (src_root && threadIdx.x < 10)
    reinterpret_cast<uint4*>(dst + 1)[threadIdx.x] = reinterpret_cast<const uint4*>(src + src_root)[threadIdx.x];
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
    // the live subtree holds at most N_root blocks, so the fresh half has room until the next move
    S->root_block = src_root ? 1u : 0u;
    S->half = half ^ 1u;
    S->n_alloc = sh_next;
    S->path_len = 0u;
    S->state = ST_CONTINUE;
    atomicAdd(&D.g->compacted_blocks, (unsigned long long)(sh_next - 1u));
    atomicAdd(&D.g->compactions, 1ull);
  }
  __syncthreads();
}

// Close the tick (one thread, after every game of the tick is done): batch size, counters, status
// for the host, next epoch.
__device__ __forceinline__ void close_tick(const Dev& D, uint32_t epoch, uint32_t n_movers) {
  Globals* G = D.g;
  __threadfence();
  const uint32_t rows = atomicExch(&G->rows_acc, 0u);
  const uint32_t waiting = atomicExch(&G->wait_acc, 0u);
  G->n_rows = rows;
  G->rows_total += rows;
  G->leaves_total += waiting;
  if (D.spec_cap) {
    // this tick's speculative rows, and how many the next tick may add: while the games themselves ask
    // for at most spec_thr rows the batch is topped up to spec_cap (the I/O buffers hold n_slots +
    // spec_cap rows, so every live game can still ask for its own)
    const uint32_t budget = G->spec_budget;
    const uint32_t acc = atomicExch(&G->spec_acc, 0u), ok = atomicExch(&G->spec_ok, 0u);
    G->spec_count[epoch & 1u] = acc < budget ? acc : budget;
    G->spec_total += ok;
    const uint32_t asked = rows - ok;
    const uint32_t next = asked <= D.spec_thr ? D.spec_cap - (asked < D.spec_cap ? asked : D.spec_cap) : 0u;
    G->spec_budget = next;
  }
  HostStatus* hs = D.status;
  hs->n_rows = rows;
  hs->n_finished = G->n_finished;
  hs->n_running = G->n_running;
  hs->n_movers = n_movers;
  hs->error = G->error;
  __threadfence_system();
  hs->tick = epoch;      // written last: the host spins on it
  G->tick = epoch + 1u;  // open the next tick
  G->n_movers = 0u;
  G->mover_claim = 0u;
  G->tail_pending = 0u;
}

// ------------------------------------------------------------------------------------------------
// K_step: the tick of every game.
// ------------------------------------------------------------------------------------------------
// NOISE = Dirichlet noise on root priors (c4a0_config.dirichlet_*): its own instantiation, so that the default
// kernel carries none of it (the extra live values cost 5 % of the tick through spills when merely compiled in)
template <bool NOISE>
__global__ void __launch_bounds__(STEP_THREADS, 7) k_step(Dev D) {  // <= 72 registers: 16,384 games in one wave
  // Launched programmatically dependent by the native loop (c4a0_engine_run_net): the CTAs may already be
  // resident while the network kernel before them drains; nothing is read before that kernel has completed.
  // (Both instructions are no-ops in an ordinary launch.)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const Lanes L = make_lanes();
  const uint32_t slot = (blockIdx.x * blockDim.x + threadIdx.x) >> 3;
  const bool valid = slot < D.n_slots;
  const bool prof = D.dbg != nullptr;
  const long long t0 = prof ? clock64() : 0;
  Game G;
  if (valid) {
    load_game(D, L, slot, G);
  } else {
    memset(&G, 0, sizeof(G));
  }
  const uint32_t st = G.state;
  const bool live = st == ST_WAIT_NN || st == ST_CONTINUE;
  const uint32_t epoch = D.g->tick;
  if (st == ST_NEED_MOVE && L.l == 0) push_mover(D, slot, epoch);  // asked for compaction after a compaction (cannot happen, kept for safety)
  const uint32_t spec_budget = D.spec_cap ? D.g->spec_budget : 0u;
  // simulations a game may run in this tick without a network row (terminal leaves, evaluation-cache hits)
  const uint32_t max_inline = spec_budget ? D.max_inline_spec : D.max_inline;
  if (__any_sync(FULL, live)) {  // (no early return: every warp takes part in closing the tick below)
  const long long t1 = prof ? clock64() : 0;
  const bool waiting = live && st == ST_WAIT_NN;
  // the row holding this game's answer was drawn by the leader of its key during the last tick
  float x = 0.0f, vq = 0.0f, vn = 0.0f;
  if (waiting) {
    const uint32_t row = (uint32_t)D.rowtag[(size_t)((epoch - 1u) & 1u) * D.n_slots + G.nn_leader];  // published last tick
    x = D.logits[(size_t)row * 7 + (L.l < 7 ? L.l : 6)];
    vq = D.qp[row];
    vn = D.qn[row];
    if (G.cache_own) {  // the evaluation-cache entry this game claimed when it asked: now readable from the next tick on
      EvalEntry* E = D.cache + (G.cache_own - 1u);
      if (L.l < 7)
        E->logit[L.l] = x;
      else
        *reinterpret_cast<float2*>(&E->qp) = make_float2(vq, vn);
    }
  }
  G.cache_own = 0u;
  const long long t2 = prof ? clock64() : 0;
  const uint32_t ns = run_games<NOISE>(D, L, G, live, st, epoch, spec_budget, max_inline, waiting, x, vq, vn);
  const long long t3 = prof ? clock64() : 0;
  const unsigned nwait = __popc(__ballot_sync(FULL, live && L.l == 0 && ns == ST_WAIT_NN));
  if ((threadIdx.x & 31) == 0 && nwait) atomicAdd(&D.g->wait_acc, nwait);
  if (D.cache != nullptr) {
    const unsigned nh = __reduce_add_sync(FULL, L.l == 0 ? G.hits : 0u), nc = __reduce_add_sync(FULL, L.l == 0 ? G.claims : 0u);
    if ((threadIdx.x & 31) == 0 && nh) atomicAdd(&D.g->cache_hits, (unsigned long long)nh);
    if ((threadIdx.x & 31) == 0 && nc) atomicAdd(&D.g->cache_inserts, (unsigned long long)nc);
  }
  __syncwarp();  // every lane's tree writes precede lane 0's publication of the game (push_mover / publish_leaf)
  if (live && L.l == 0) {
    store_game(D, G, ns);
    if (ns == ST_NEED_MOVE) push_mover(D, slot, epoch);
    if (ns == ST_WAIT_NN)
      publish_leaf(D, slot, G.leaf.mask, G.leaf.value, stored_leaf_model(D, G), epoch, G.cache_own);
    if (prof) {  // cycles of this game's warp per phase (c4a0_engine_debug_phases)
      uint32_t* o = D.dbg + (size_t)slot * 8;
      o[0] = (uint32_t)(t1 - t0);         // load slot state
      o[1] = (uint32_t)(t2 - t1);         // fetch the network's answer
      o[2] = G.sims;                      // simulations this tick
      o[3] = (uint32_t)(t3 - t2);         // softmax/expand/backup, moves, selection passes, in-kernel simulations
      o[4] = G.term + G.hits;             // ... of which needed no network row (terminal leaf, cache hit)
      o[5] = G.len;                       // depth of the selected leaf
      o[6] = (uint32_t)(clock64() - t3);  // store
      o[7] = (uint32_t)(clock64() - t0);  // whole tick
    }
  }
  }
  if (D.spec_cap) spec_collect(D, epoch);
  // ---- compactions: a CTA that is done with its own games takes arenas off the list while slower CTAs are
  // still running, so the copies overlap the tail of the tick instead of following it -------------------
  __shared__ uint32_t sh_last, sh_take;
  __syncthreads();
  for (;;) {
    if (threadIdx.x == 0) {
      uint32_t got = 0xffffffffu;
      uint32_t c = *reinterpret_cast<volatile uint32_t*>(&D.g->mover_claim);
      for (;;) {
        const uint32_t n = *reinterpret_cast<volatile uint32_t*>(&D.g->n_movers);
        // a long backlog (minimal arenas: every re-root compacts) is left to k_tail, whose larger CTAs copy an
        // arena faster and do not hold up the later waves of this kernel (config 4: 131,072 games, 8 waves)
        if (c >= n || n - c > STEAL_BACKLOG_MAX) break;
        const uint32_t prev = atomicCAS(&D.g->mover_claim, c, c + 1u);
        if (prev == c) {
          got = mover_slot(D, c, epoch);
          break;
        }
        c = prev;
      }
      sh_take = got;
    }
    __syncthreads();
    const uint32_t take = sh_take;
    if (take == 0xffffffffu) break;
    compact_one(D, take);  // ends with a CTA barrier: sh_take can be rewritten
  }
  // ---- the last CTA to finish closes the tick (or hands it to k_tail) ---------------------------
  if (threadIdx.x == 0) {
    __threadfence();
    sh_last = atomicAdd(&D.g->step_done, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    const uint32_t n_movers = *reinterpret_cast<volatile uint32_t*>(&D.g->n_movers);
    const uint32_t taken = *reinterpret_cast<volatile uint32_t*>(&D.g->mover_claim);  // final: every CTA has left the loop above
    if (threadIdx.x == 0) D.g->step_done = 0u;
    if (n_movers - taken <= INLINE_COMPACTIONS) {
      for (uint32_t m = taken; m < n_movers; m++) compact_one(D, mover_slot(D, m, epoch));
      if (threadIdx.x == 0) close_tick(D, epoch, n_movers);
    } else if (threadIdx.x == 0) {  // a late burst of compactions: one CTA per arena in k_tail
      D.g->tail_pending = 1u;
      __threadfence_system();
      D.status->need_tail = epoch;
    }
  }
}

__global__ void k_ln_table(float* tab, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) tab[i] = c4::c4_logf((float)i);
}

// Seat the first min(n_slots, n_req) games (self_play.rs:55-58).
__global__ void k_init_globals(Dev D, uint32_t n_req) {  // runs alone, before k_init
  Globals z;
  memset(&z, 0, sizeof(z));
  z.tick = 1u;
  z.n_req = n_req;
  z.next_req = n_req < D.n_slots ? n_req : D.n_slots;
  z.n_running = z.next_req;
  z.tail_pending = 1u;  // k_tail closes the seating 'tick'
  *D.g = z;
}

__global__ void k_init(Dev D, uint32_t n_req) {
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= D.n_slots) return;
  Slot S;
  memset(&S, 0, sizeof(S));
  S.n_alloc = 1u;
  if (slot < n_req) {
    S.req = slot;
    S.state = ST_WAIT_NN;
    S.model0 = D.p0[slot];
    S.model1 = D.p1[slot];
  }
  D.slots[slot] = S;
  if (slot < n_req) {
    atomicAdd(&D.g->wait_acc, 1u);
    publish_leaf(D, slot, 0ull, 0ull, S.model0, 1u, 0u);
  }
}

// ------------------------------------------------------------------------------------------------
// K_tail: only does something when k_step left the tick open (many arenas filled up at once, or
// right after k_init): one CTA per arena to compact, the last CTA to finish closes the tick.
// ------------------------------------------------------------------------------------------------
constexpr int TAIL_THREADS = 256;
constexpr int TAIL_CTAS = 148;

__global__ void __launch_bounds__(TAIL_THREADS) k_tail(Dev D) {
  __shared__ uint32_t sh_last;
  if (*reinterpret_cast<volatile uint32_t*>(&D.g->tail_pending) == 0u) return;  // grid-uniform
  const uint32_t epoch = D.g->tick;
  const uint32_t n_movers = D.g->n_movers, taken = D.g->mover_claim;  // k_step's CTAs compacted entries [0, taken)
  for (uint32_t m = taken + blockIdx.x; m < n_movers; m += gridDim.x) compact_one(D, mover_slot(D, m, epoch));
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    sh_last = atomicAdd(&D.g->tail_done, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last && threadIdx.x == 0) {  // every CTA (and, by stream order, all of k_step) is done
    D.g->tail_done = 0u;
    close_tick(D, epoch, n_movers);
  }
}

// ------------------------------------------------------------------------------------------------
// Synthetic evaluators (parity tiers E0 / E1) and row views
// ------------------------------------------------------------------------------------------------
__global__ void k_eval_builtin(Dev D, int kind, float* logits, float* qp, float* qn) {
  uint32_t row = blockIdx.x * blockDim.x + threadIdx.x;
  // rows of the batch: n_rows once the tick is closed; while a burst of compactions keeps it open for
  // k_tail (the native loop has already enqueued the network behind k_step) the row counter itself
  const uint32_t closed = D.g->n_rows, open = *reinterpret_cast<volatile uint32_t*>(&D.g->rows_acc);
  if (row >= (closed > open ? closed : open)) return;
  if (kind == C4A0_EVAL_UNIFORM) {
    for (int k = 0; k < 7; k++) logits[(size_t)row * 7 + k] = 0.0f;
    qp[row] = 0.0f;
    qn[row] = 0.0f;
    return;
  }
  uint64_t mask = D.row_mask[row], value = D.row_value[row], model = D.row_model[row];
  uint64_t h = splitmix64(mask * 0x9E3779B97F4A7C15ULL ^ splitmix64(value ^ model));
  for (int k = 0; k < 7; k++) {
    uint64_t hk = splitmix64(h + (uint64_t)k);
    const float lg = (float)(uint32_t)(hk >> 48) * (1.0f / 8192.0f) - 4.0f;
    logits[(size_t)row * 7 + k] = kind == C4A0_EVAL_HASH_FLAT ? lg * 0.0625f : lg;
  }
  uint64_t hq = splitmix64(h + 7);
  const float qs = kind == C4A0_EVAL_HASH_FLAT ? 0.1f : 0.75f;
  qp[row] = ((float)(uint32_t)((hq >> 48) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * qs;
  qn[row] = ((float)(uint32_t)((hq >> 32) & 0xffff) * (1.0f / 32768.0f) - 1.0f) * qs;
}

// Training tensors straight from the sample store (what src/c4a0/training.py:317-333 builds with a
// Python loop of Sample.to_numpy() and Sample.flip_h()): sample k of game g goes to row
// offsets[g - first] + k; with `flip` its mirror image (c4r.rs:289-299, types.rs:115-122) goes to row
// total + offsets[g - first] + k.  One thread per sample.
__global__ void k_export(Dev D, uint32_t first, uint32_t n, const uint32_t* offsets, uint32_t total, int flip,
                         float* pos, float* policy, float* qp, float* qn) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t gi = t / MAXS, k = t % MAXS;
  if (gi >= n) return;
  const uint32_t g = first + gi;
  if (k >= D.n_samples[g]) return;
  const size_t si = (size_t)g * MAXS + k;
  Pos p{D.s_mask[si], D.s_value[si]};
  float pol[7];
#pragma unroll
  for (int i = 0; i < 7; i++) pol[i] = D.s_policy[si * 7 + i];
  const float a = D.s_qp[si], b = D.s_qn[si];
  for (int rep = 0; rep < (flip ? 2 : 1); rep++) {
    const size_t row = (size_t)offsets[gi] + k + (rep ? total : 0u);
    const Pos q = rep ? c4::flip_h(p) : p;
    float4* dst = reinterpret_cast<float4*>(pos + row * 84);
    const uint64_t mine = q.mask & q.value, theirs = q.mask & ~q.value;
    const uint64_t lo = mine | (theirs << 42), hi = theirs >> 22;
#pragma unroll
    for (int v = 0; v < 21; v++) {
      uint32_t bits = (uint32_t)((v < 16 ? lo >> (4 * v) : hi >> (4 * (v - 16))) & 0xfull);
      dst[v] = make_float4((float)(bits & 1u), (float)((bits >> 1) & 1u), (float)((bits >> 2) & 1u), (float)((bits >> 3) & 1u));
    }
#pragma unroll
    for (int i = 0; i < 7; i++) policy[row * 7 + i] = rep ? pol[6 - i] : pol[i];
    qp[row] = a;
    qn[row] = b;
  }
}

__global__ void k_sum_counters(Dev D, unsigned long long* out4) {
  unsigned long long a = 0, b = 0, c = 0, d = 0;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < D.n_slots; i += gridDim.x * blockDim.x) {
    const Slot* S = D.slots + i;
    a += S->c_sims;
    b += S->c_exp;
    c += S->c_term;
    d += S->c_depth;
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
  Globals* h_globals = nullptr;   // pinned
  HostStatus* h_status = nullptr; // pinned + mapped
  float *b_logits = nullptr, *b_qp = nullptr, *b_qn = nullptr;  // writable aliases for eval_builtin
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
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

int launch_post(c4a0_engine* e, cudaStream_t s) {
  k_tail<<<TAIL_CTAS, TAIL_THREADS, 0, s>>>(e->D);
  CK(cudaGetLastError());
  return 0;
}

// Enqueue one tick.  `ev` (4 events) brackets k_step and k_tail when given (ev[1] == ev[2]).  k_step
// closes the tick by itself unless a burst of arenas needs compaction; with_tail enqueues k_tail
// unconditionally (it returns at once when there is nothing to do) so that callers that cannot look
// at the status word in between (engine_step(), CUDA-graph capture) always get a closed tick.
int launch_tick(c4a0_engine* e, cudaStream_t s, cudaEvent_t* ev, bool with_tail, bool pdl = false) {
  const Dev& D = e->D;
  if (ev) CK(cudaEventRecord(ev[0], s));
  if (pdl) {  // programmatic dependent launch behind the network kernel (see k_step)
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(blocks_for((size_t)D.n_slots * 8, STEP_THREADS));
    cfg.blockDim = dim3(STEP_THREADS);
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    CK(D.dir_eps > 0.0f ? cudaLaunchKernelEx(&cfg, k_step<true>, D) : cudaLaunchKernelEx(&cfg, k_step<false>, D));
  } else if (D.dir_eps > 0.0f) {
    k_step<true><<<blocks_for((size_t)D.n_slots * 8, STEP_THREADS), STEP_THREADS, 0, s>>>(D);
  } else {
    k_step<false><<<blocks_for((size_t)D.n_slots * 8, STEP_THREADS), STEP_THREADS, 0, s>>>(D);
  }
  if (ev) CK(cudaEventRecord(ev[1], s));
  if (ev) CK(cudaEventRecord(ev[2], s));
  if (with_tail) {
    int r = launch_post(e, s);
    if (r) return r;
  }
  if (ev) CK(cudaEventRecord(ev[3], s));
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
  if (cfg->plane_stride && (cfg->plane_stride < 84 || cfg->plane_stride % 4 ||
                            (cfg->plane_stride > 84 && cfg->plane_stride < 88)))
    return fail(C4A0_E_INVALID, "plane_stride must be 0, 84, or a multiple of 4 that is >= 88");
  if (cfg->n_mcts_iterations > (1u << 24))
    return fail(C4A0_E_INVALID, "n_mcts_iterations above 2^24 is not exactly representable in f32");
  if (cfg->n_slots > (1u << 24)) return fail(C4A0_E_INVALID, "n_slots above 2^24 is not supported");
  if (cfg->arena_blocks && cfg->arena_blocks < cfg->n_mcts_iterations + 2)
    return fail(C4A0_E_INVALID, "arena_blocks must be 0 or >= n_mcts_iterations + 2");
  if (cfg->arena_blocks >= (1u << 29)) return fail(C4A0_E_INVALID, "arena_blocks too large");
  if ((cfg->flags & C4A0_FLAG_SPECULATE) && !(cfg->flags & C4A0_FLAG_EVAL_CACHE))
    return fail(C4A0_E_INVALID, "C4A0_FLAG_SPECULATE needs C4A0_FLAG_EVAL_CACHE");
  if (!(cfg->dirichlet_alpha >= 0.0f) || !(cfg->dirichlet_epsilon >= 0.0f) || cfg->dirichlet_epsilon > 1.0f)
    return fail(C4A0_E_INVALID, "dirichlet_alpha must be >= 0 and dirichlet_epsilon in [0, 1]");
  int r = c4host::no_gpu_error();
  if (r) return r;
  CK(cudaSetDevice(cfg->device));
  c4a0_engine* e = new c4a0_engine();
  e->cfg = *cfg;
  Dev& D = e->D;
  memset(&D, 0, sizeof(D));
  D.n_slots = cfg->n_slots;
  D.n_iter = cfg->n_mcts_iterations;
  // index 0 is unused; a tree holds at most n_iter expanded nodes, so n_iter + 2 always suffices
  D.cap = cfg->arena_blocks ? cfg->arena_blocks : cfg->n_mcts_iterations + 2;
  const bool use_cache = (cfg->flags & C4A0_FLAG_EVAL_CACHE) != 0;
  D.max_inline = cfg->max_inline_sims ? cfg->max_inline_sims : (use_cache ? 3 : 2);
  D.plane_bf16 = cfg->plane_dtype == C4A0_PLANES_BF16;
  D.plane_stride = cfg->plane_stride ? cfg->plane_stride : 84;
  D.dedup = (cfg->flags & C4A0_FLAG_NO_DEDUP) ? 0u : 1u;
  D.c_expl = cfg->c_exploration;
  D.c_pen = cfg->c_ply_penalty;
  D.dir_alpha = cfg->dirichlet_alpha;
  D.dir_eps = cfg->dirichlet_alpha > 0.0f ? cfg->dirichlet_epsilon : 0.0f;
  size_t S = cfg->n_slots, R = cfg->max_requests;
  size_t T = 1;
  while (T < 2 * S) T <<= 1;
  D.table_mask = (uint32_t)(T - 1);
  if (use_cache && (cfg->flags & C4A0_FLAG_SPECULATE)) {
    D.spec_cap = cfg->spec_rows ? cfg->spec_rows : 8192u;
    if (D.spec_cap > cfg->n_slots) D.spec_cap = cfg->n_slots;
    // speculate while the games themselves ask for at most this many rows: beyond that the spare rows
    // cover too small a part of what the live games will want next (measured on the bench job)
    D.spec_thr = D.spec_cap / 2 < 1024u ? D.spec_cap / 2 : 1024u;
    if (const char* env = getenv("C4A0_SPEC_THR")) D.spec_thr = (uint32_t)atoi(env);  // tuning knob (results do not depend on it)
    D.max_inline_spec = 4 * D.max_inline;
    if (const char* env = getenv("C4A0_INLINE_SPEC_MULT")) D.max_inline_spec = (uint32_t)atoi(env) * D.max_inline;
  }

  D.row_cap = cfg->n_slots + D.spec_cap;
  const size_t RC = D.row_cap;
  DA(D.slots, S); DA(D.path, S * PATH_STRIDE); DA(D.row_slot, RC); DA(D.row_model, RC); DA(D.rowtag, 2 * S);
  DA(D.blocks, S * 2 * (size_t)D.cap);
  DA(D.table, T);
  uint64_t *gid, *p0, *p1;
  DA(gid, R); DA(p0, R); DA(p1, R);
  D.game_id = gid; D.p0 = p0; D.p1 = p1;
  DA(D.n_samples, R); DA(D.s_mask, R * MAXS); DA(D.s_value, R * MAXS);
  DA(D.s_policy, R * MAXS * 7); DA(D.s_qp, R * MAXS); DA(D.s_qn, R * MAXS);
  DA(D.g, 1); DA(D.movers, S);
  DA(e->scratch4, 4); DA(D.row_mask, RC); DA(D.row_value, RC);
  size_t CE = 0;
  if (use_cache) {
    // default: room for 16 entries per simulation of one move of every resident game (a job evaluates a
    // few times that many distinct positions; later answers replace earlier ones), at most a quarter of
    // the memory that is still free.  Bench job: 2^28 entries = 17 GB; 2^27 costs 4.5 % of the step (a
    // fuller table drops more speculative claims and delays more replacements), 2^29 gains another 2 %
    size_t want = cfg->eval_cache_entries ? cfg->eval_cache_entries : S * (size_t)cfg->n_mcts_iterations * 16;
    if (want < 1024) want = cfg->eval_cache_entries ? (want < 2 ? 2 : want) : 1024;
    CE = 1;
    while (CE < want && CE < ((size_t)1 << 31)) CE <<= 1;
    if (!cfg->eval_cache_entries) {
      size_t free_b = 0, total_b = 0;
      if (cudaMemGetInfo(&free_b, &total_b) != cudaSuccess) {
        c4a0_engine_destroy(e);
        return fail(C4A0_E_CUDA, "cudaMemGetInfo failed");
      }
      while (CE > 1024 && CE * sizeof(EvalEntry) > free_b / 4) CE >>= 1;
    }
    D.cache_mask = (uint32_t)(CE - 1);
    DA(D.cache, CE);
  }
  {  // a node is selected through with at most n_iter visits
    float* tab = nullptr;
    D.ln_n = cfg->n_mcts_iterations < (1u << 20) ? cfg->n_mcts_iterations + 2 : (1u << 20);
    DA(tab, D.ln_n);
    k_ln_table<<<blocks_for(D.ln_n, 256), 256>>>(tab, D.ln_n);
    D.ln_tab = tab;
  }
  if (D.spec_cap) DA(D.spec_list, 2 * (size_t)D.spec_cap);
  cudaError_t err = cudaGetLastError();
  if (err == cudaSuccess) err = cudaMemset(D.slots, 0, S * sizeof(Slot));
  if (err == cudaSuccess && CE) err = cudaMemset(D.cache, 0, CE * sizeof(EvalEntry));
  if (err == cudaSuccess) err = cudaMemset(D.g, 0, sizeof(Globals));
  if (err == cudaSuccess) err = cudaMemset(D.table, 0, T * sizeof(unsigned long long));
  if (err == cudaSuccess) err = cudaMallocHost((void**)&e->h_globals, sizeof(Globals));
  if (err == cudaSuccess) err = cudaHostAlloc((void**)&e->h_status, sizeof(HostStatus), cudaHostAllocMapped);
  if (err == cudaSuccess) {
    memset((void*)e->h_status, 0, sizeof(HostStatus));
    err = cudaHostGetDevicePointer((void**)&D.status, (void*)e->h_status, 0);
  }
  if (err == cudaSuccess) err = cudaDeviceSynchronize();  // the fills above ran on the default stream
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
uint32_t c4a0_engine_io_rows(const c4a0_engine* e) { return e ? e->D.row_cap : 0; }

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
  CK(cudaMemsetAsync(D.rowtag, 0, 2 * (size_t)D.n_slots * sizeof(unsigned long long), s));
  CK(cudaMemsetAsync(D.movers, 0, (size_t)D.n_slots * sizeof(unsigned long long), s));  // epochs restart at 1
  D.job = (D.job + 1u) & 0x7fffffffu;  // entries of the evaluation cache written by earlier jobs become stale
  if (D.job == 0u) {                   // (after 2^31 jobs: start over with an empty table)
    D.job = 1u;
    if (D.cache) CK(cudaMemsetAsync(D.cache, 0, ((size_t)D.cache_mask + 1) * sizeof(EvalEntry), s));
  }
  k_init_globals<<<1, 1, 0, s>>>(D, n);
  k_init<<<blocks_for(D.n_slots, 256), 256, 0, s>>>(D, n);
  CK(cudaGetLastError());
  int r = launch_post(e, s);
  if (r) return r;
  CK(cudaStreamSynchronize(s));  // the host arrays may be freed by the caller after return
  e->n_req = n;
  e->have_requests = true;
  e->steps = 0;
  return 0;
}

int c4a0_engine_step(c4a0_engine* e, void* stream) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (!e->have_requests) return fail(C4A0_E_INVALID, "set_requests() must precede step()");
  return launch_tick(e, (cudaStream_t)stream, nullptr, true);
}

int c4a0_engine_step_timed(c4a0_engine* e, void* stream, float* ms_step, float* ms_tail) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (!e->have_requests) return fail(C4A0_E_INVALID, "set_requests() must precede step()");
  cudaStream_t s = (cudaStream_t)stream;
  if (!e->ev[0])
    for (int i = 0; i < 4; i++) CK(cudaEventCreate(&e->ev[i]));
  int r = launch_tick(e, s, e->ev, true);
  if (r) return r;
  CK(cudaStreamSynchronize(s));
  float a = 0, b = 0;
  CK(cudaEventElapsedTime(&a, e->ev[0], e->ev[1]));
  CK(cudaEventElapsedTime(&b, e->ev[2], e->ev[3]));
  if (ms_step) *ms_step = a;
  if (ms_tail) *ms_tail = b;
  return 0;
}

int c4a0_engine_debug_phases(c4a0_engine* e, void* stream, uint32_t* out8_per_slot) {
  if (!e || !out8_per_slot) return fail(C4A0_E_INVALID, "null argument");
  if (!e->have_requests) return fail(C4A0_E_INVALID, "set_requests() must precede step()");
  cudaStream_t s = (cudaStream_t)stream;
  size_t n = (size_t)e->D.n_slots * 8;
  uint32_t* d = nullptr;
  CK(cudaMalloc((void**)&d, n * 4));
  CK(cudaMemsetAsync(d, 0, n * 4, s));
  e->D.dbg = d;
  int r = launch_tick(e, s, nullptr, true);
  e->D.dbg = nullptr;
  if (!r) {
    cudaError_t err = cudaMemcpyAsync(out8_per_slot, d, n * 4, cudaMemcpyDeviceToHost, s);
    if (err == cudaSuccess) err = cudaStreamSynchronize(s);
    if (err != cudaSuccess) r = fail(C4A0_E_CUDA, "debug copy failed: %s", cudaGetErrorString(err));
  }
  cudaFree(d);
  return r;
}

int c4a0_engine_eval_builtin(c4a0_engine* e, int kind, void* stream) {
  if (!e || !e->io_bound) return fail(C4A0_E_INVALID, "engine not bound");
  if (kind != C4A0_EVAL_UNIFORM && kind != C4A0_EVAL_HASH && kind != C4A0_EVAL_HASH_FLAT) return fail(C4A0_E_INVALID, "bad evaluator kind");
  k_eval_builtin<<<blocks_for(e->D.row_cap, 256), 256, 0, (cudaStream_t)stream>>>(e->D, kind, e->b_logits, e->b_qp, e->b_qn);
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
  out->n_movers = e->h_status->n_movers;
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
  out->compactions = g.compactions;
  out->cache_hits = g.cache_hits;
  out->cache_inserts = g.cache_inserts;
  out->spec_rows = g.spec_total;
  return 0;
}

int c4a0_engine_fetch_rows(c4a0_engine* e, uint32_t* n_rows, uint64_t* leaf_mask, uint64_t* leaf_value,
                           uint64_t* model_id, void* stream) {
  if (!e || !n_rows) return fail(C4A0_E_INVALID, "null argument");
  cudaStream_t s = (cudaStream_t)stream;
  CK(cudaMemcpyAsync(e->h_globals, e->D.g, sizeof(Globals), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  uint32_t n = e->h_globals->n_rows;
  *n_rows = n;
  if (n && leaf_mask) CK(cudaMemcpyAsync(leaf_mask, e->D.row_mask, n * 8, cudaMemcpyDeviceToHost, s));
  if (n && leaf_value) CK(cudaMemcpyAsync(leaf_value, e->D.row_value, n * 8, cudaMemcpyDeviceToHost, s));
  if (n && model_id) CK(cudaMemcpyAsync(model_id, e->D.row_model, n * 8, cudaMemcpyDeviceToHost, s));
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

int c4a0_engine_export_samples(c4a0_engine* e, uint32_t first, uint32_t n, const uint32_t* offsets_dev,
                               uint32_t total, int flip, float* pos_dev, float* policy_dev, float* qp_dev,
                               float* qn_dev, void* stream) {
  if (!e || !offsets_dev || !pos_dev || !policy_dev || !qp_dev || !qn_dev) return fail(C4A0_E_INVALID, "null argument");
  if ((uint64_t)first + n > e->n_req) return fail(C4A0_E_INVALID, "sample range out of bounds");
  if (((uintptr_t)pos_dev & 15u) != 0) return fail(C4A0_E_INVALID, "pos_dev must be 16-byte aligned");
  if (n == 0) return 0;
  k_export<<<blocks_for((size_t)n * MAXS, 256), 256, 0, (cudaStream_t)stream>>>(e->D, first, n, offsets_dev, total, flip,
                                                                             pos_dev, policy_dev, qp_dev, qn_dev);
  CK(cudaGetLastError());
  return 0;
}

int c4a0_engine_rows_dev(c4a0_engine* e, uint32_t** row_slot_dev, uint64_t** row_model_dev) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (row_slot_dev) *row_slot_dev = e->D.row_slot;
  if (row_model_dev) *row_model_dev = e->D.row_model;
  return 0;
}

int c4a0_engine_rows_count_dev(c4a0_engine* e, const uint32_t** closed_dev, const uint32_t** open_dev) {
  if (!e) return fail(C4A0_E_INVALID, "null engine");
  if (closed_dev) *closed_dev = &e->D.g->n_rows;
  if (open_dev) *open_dev = &e->D.g->rows_acc;
  return 0;
}

int c4a0_engine_slot_info(c4a0_engine* e, uint32_t slot, c4a0_slot_info* out, void* stream) {
  if (!e || !out) return fail(C4A0_E_INVALID, "null argument");
  if (slot >= e->D.n_slots) return fail(C4A0_E_INVALID, "slot out of range");
  cudaStream_t s = (cudaStream_t)stream;
  Slot S;
  CK(cudaMemcpyAsync(&S, e->D.slots + slot, sizeof(Slot), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  out->state = S.state;
  out->request = S.req;
  out->n_moves = S.n_moves;
  out->root_visits = S.root_N;
  out->root_mask = S.root_mask;
  out->root_value = S.root_value;
  out->root_q_sum_penalty = S.root_Qp;
  out->root_q_sum_no_penalty = S.root_Qn;
  out->n_blocks = S.n_alloc ? S.n_alloc - 1 : 0;
  unsigned long long tag = 0;
  if (S.state == ST_WAIT_NN && S.nn_leader < e->D.n_slots) {
    uint32_t tick = 0;  // the open tick; the leaf was published in the one before
    CK(cudaMemcpyAsync(&tick, &e->D.g->tick, 4, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
    CK(cudaMemcpyAsync(&tag, e->D.rowtag + (size_t)((tick - 1u) & 1u) * e->D.n_slots + S.nn_leader, 8, cudaMemcpyDeviceToHost, s));
    CK(cudaStreamSynchronize(s));
  }
  out->nn_row = (uint32_t)tag;
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
      put(B.rec[c].N);
      put(c4::f32_bits(B.rec[c].Qp));
      put(c4::f32_bits(B.rec[c].Qn));
      put(c4::f32_bits(B.rec[c].P));
      if (B.child[c]) children(B.child[c], c4::make_move(pos, c));
    }
  }
};
}  // namespace

int c4a0_engine_dump_tree(c4a0_engine* e, uint32_t slot, uint32_t* buf, size_t cap, size_t* needed,
                          void* stream) {
  if (!e || !needed) return fail(C4A0_E_INVALID, "null argument");
  if (slot >= e->D.n_slots) return fail(C4A0_E_INVALID, "slot out of range");
  cudaStream_t s = (cudaStream_t)stream;
  const Dev& D = e->D;
  Slot S;
  CK(cudaMemcpyAsync(&S, D.slots + slot, sizeof(Slot), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  std::vector<Block> host(S.n_alloc ? S.n_alloc : 1);
  const Block* src = D.blocks + ((size_t)slot * 2 + S.half) * D.cap;
  CK(cudaMemcpyAsync(host.data(), src, host.size() * sizeof(Block), cudaMemcpyDeviceToHost, s));
  CK(cudaStreamSynchronize(s));
  Dumper d{host.data(), buf, buf ? cap : 0, 0};
  d.put(S.root_block ? 2u : 1u);
  d.put(S.root_N);
  d.put(c4::f32_bits(S.root_Qp));
  d.put(c4::f32_bits(S.root_Qn));
  if (S.root_block) d.children(S.root_block, Pos{S.root_mask, S.root_value});
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
}  // extern "C"

namespace {
// graphs != nullptr: the evaluator of engine i is graphs[i][0..n_graphs[i]) (CUDA graphs, one per batch-size bucket);
// nets != nullptr: it is nets[i], the library's own network kernel, launched directly — programmatically dependent
// on the tree tick before it, and the next tick on it — unless C4A0_PDL=0.
int run_loop(c4a0_engine* const* engines, uint32_t n_engines, const c4a0_nn_graph* const* graphs, const uint32_t* n_graphs,
             c4a0_net* const* nets, void* const* streams, uint64_t max_ticks, uint32_t time_kernels_every,
             c4a0_run_report* out) {
  if (!engines || (!graphs && !nets) || (graphs && !n_graphs) || !streams || !out || n_engines == 0 || n_engines > 8)
    return fail(C4A0_E_INVALID, "bad argument");
  memset(out, 0, sizeof(*out));
  bool pdl = nets != nullptr;
  if (const char* env = getenv("C4A0_PDL")) pdl = pdl && env[0] != '0';
  struct Lane {
    c4a0_engine* e;
    cudaStream_t s;
    uint32_t expect;     // status tick the host waits for next
    uint32_t spec_rows;  // rows covered by the network graph already enqueued behind that tick
    uint32_t prev_rows;  // rows of the tick before the last closed one
    bool done;
    cudaEvent_t t0, t1;
  };
  std::vector<Lane> lanes(n_engines);
  std::vector<cudaEvent_t> kev;  // three per sampled tick: before k_step, after k_step, after the network
  for (uint32_t i = 0; i < n_engines; i++) {
    c4a0_engine* e = engines[i];
    if (!e || !e->have_requests) return fail(C4A0_E_INVALID, "engine %u has no requests", i);
    if (nets) {
      if (!nets[i]) return fail(C4A0_E_INVALID, "engine %u has no network", i);
    } else {
      if (n_graphs[i] == 0 || !graphs[i]) return fail(C4A0_E_INVALID, "engine %u has no network graphs", i);
      for (uint32_t k = 0; k < n_graphs[i]; k++) {
        if (!graphs[i][k].graph_exec) return fail(C4A0_E_INVALID, "null graph_exec");
        if (k && graphs[i][k].rows <= graphs[i][k - 1].rows) return fail(C4A0_E_INVALID, "network graphs must be sorted by rows");
      }
      if (graphs[i][n_graphs[i] - 1].rows < e->D.row_cap)
        return fail(C4A0_E_INVALID, "the largest network graph must cover c4a0_engine_io_rows() rows");
    }
    lanes[i] = Lane{e, (cudaStream_t)streams[i], e->h_status->tick, 0u, 0u, e->n_req == 0, nullptr, nullptr};
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
  nvtxRangePushA("c4a0_engine_run");
  // guess for the next tick's rows = this tick's rows * (1 + 2^-spec_shift): rows drift slowly
  uint32_t spec_shift = 8;
  if (const char* env = getenv("C4A0_SPEC_SHIFT")) spec_shift = (uint32_t)atoi(env) & 31u;
  // the network on rows [0, >= rows): the smallest captured graph that covers them
  auto launch_nn = [&](uint32_t i, uint32_t rows, uint32_t* covered) -> int {
    Lane& L = lanes[i];
    if (nets) {  // one kernel whatever the batch: it reads the tick's row count on the device
      int r = c4a0_net_forward_ex(nets[i], L.e->D.row_cap, L.s, pdl ? C4A0_NET_LAUNCH_PDL : 0u);
      if (r) return r;
      out->nn_launches++;
      out->bucket_launches[0]++;
      out->nn_rows_launched += L.e->D.row_cap;
      if (covered) *covered = L.e->D.row_cap;
      return 0;
    }
    const c4a0_nn_graph* g = graphs[i];
    uint32_t k = 0;
    while (k + 1 < n_graphs[i] && g[k].rows < rows) k++;
    CK(cudaGraphLaunch((cudaGraphExec_t)g[k].graph_exec, L.s));
    out->nn_launches++;
    out->bucket_launches[k < 31 ? k : 31]++;
    out->nn_rows_launched += g[k].rows;
    if (covered) *covered = g[k].rows;
    return 0;
  };
  // One round for engine i whose last closed tick packed `rows` rows (and whose network answers for
  // them are in flight or done): the next tick, then — without waiting for that tick's row count —
  // the network for it, sized from the current count plus a margin.  The host round trip (status
  // over PCIe, cudaGraphLaunch) then overlaps the network instead of idling the GPU; if the tick
  // turns out to need more rows than guessed, the network is simply run again at the right size.
  auto launch_round = [&](uint32_t i, uint32_t rows) -> int {
    Lane& L = lanes[i];
    cudaEvent_t* ev = nullptr;
    struct Range {  // NVTX: the host-side enqueue of one tick (k_step + the network behind it)
      Range() { nvtxRangePushA("tick: k_step + network"); }
      ~Range() { nvtxRangePop(); }
    } range;
    if (time_kernels_every && (L.e->steps % time_kernels_every) == 0 && kev.size() < 3 * 4096) {
      size_t b = kev.size();
      for (int q = 0; q < 3; q++) {
        cudaEvent_t x;
        CK(cudaEventCreate(&x));
        kev.push_back(x);
      }
      ev = &kev[b];
      CK(cudaEventRecord(ev[0], L.s));
    }
    int r = launch_tick(L.e, L.s, nullptr, false, pdl && !ev);
    if (r) return r;
    if (ev) CK(cudaEventRecord(ev[1], L.s));
    // ... or, while the batch is growing, by as much again as it grew last time
    uint32_t margin = (rows >> spec_shift) > 16 ? (rows >> spec_shift) : 16;
    if (rows > L.prev_rows && rows - L.prev_rows > margin) margin = rows - L.prev_rows;
    L.prev_rows = rows;
    const uint32_t guess = rows + margin < L.e->D.row_cap ? rows + margin : L.e->D.row_cap;
    r = launch_nn(i, guess, &L.spec_rows);
    if (r) return r;
    if (ev) CK(cudaEventRecord(ev[2], L.s));
    L.expect++;
    out->ticks++;
    return 0;
  };
  for (uint32_t i = 0; i < n_engines && !rc; i++) {
    Lane& L = lanes[i];
    CK(cudaEventRecord(L.t0, L.s));
    if (L.done) continue;
    // the planes of the initial roots are already packed (set_requests): their network first
    rc = launch_nn(i, L.e->h_status->n_rows, nullptr);
    if (!rc) rc = launch_round(i, L.e->h_status->n_rows);
  }
  uint32_t remaining = 0;
  for (auto& L : lanes) remaining += L.done ? 0 : 1;
  uint32_t cur = 0;
  while (remaining && !rc) {
    Lane& L = lanes[cur];
    if (!L.done) {
      // wait for the tick's status (written through mapped memory when the tick closes)
      uint64_t spins = 0;
      bool tail_launched = false;
      const auto wait0 = std::chrono::steady_clock::now();
      while (L.e->h_status->tick != L.expect) {
        if (!tail_launched && L.e->h_status->need_tail == L.expect) {
          // k_step left the tick open: a burst of arenas to compact -> one CTA per arena.  (k_tail
          // publishes no leaves, so the network already enqueued behind k_step stays valid.)
          std::atomic_thread_fence(std::memory_order_acquire);
          if (launch_post(L.e, L.s)) {
            rc = C4A0_E_CUDA;
            break;
          }
          tail_launched = true;
          out->tail_launches++;
          continue;
        }
        if (++spins > 2000) {
          // never spin forever: a faulted kernel, an idle stream or a stalled device ends the run
          cudaError_t q = cudaStreamQuery(L.s);
          if (q == cudaSuccess) {
            if (L.e->h_status->tick != L.expect && L.e->h_status->need_tail != L.expect) {
              rc = fail(C4A0_E_CUDA, "tick status never arrived (stream idle)");
              break;
            }
          } else if (q != cudaErrorNotReady) {
            rc = fail(C4A0_E_CUDA, "stream failed while waiting for the tick: %s", cudaGetErrorString(q));
            break;
          }
          if ((spins & 0xfff) == 0 &&
              std::chrono::duration<double>(std::chrono::steady_clock::now() - wait0).count() > 60.0) {
            rc = fail(C4A0_E_CUDA, "no tick status for 60 s");
            break;
          }
          std::this_thread::yield();
        }
      }
      out->host_wait_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wait0).count();
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
        const auto l0 = std::chrono::steady_clock::now();
        const uint32_t rows = L.e->h_status->n_rows;
        if (rows > L.spec_rows) {  // the guess was too small: evaluate again, all live rows
          rc = launch_nn(cur, rows, nullptr);
          out->nn_relaunches++;
        }
        if (!rc) rc = launch_round(cur, rows);
        out->host_launch_ms += std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - l0).count();
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
    double sum[2] = {0, 0};
    uint32_t n = 0;
    for (size_t q = 0; q + 2 < kev.size(); q += 3) {
      float d[2];
      if (cudaEventElapsedTime(&d[0], kev[q], kev[q + 1]) == cudaSuccess &&
          cudaEventElapsedTime(&d[1], kev[q + 1], kev[q + 2]) == cudaSuccess) {
        sum[0] += d[0];
        sum[1] += d[1];
        n++;
      }
    }
    out->kernel_samples = n;
    out->k_step_ms_sum = sum[0];
    out->nn_ms_sum = sum[1];
  }
  out->wall_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - wall0).count();
  nvtxRangePop();
  cleanup();
  return rc;
}
}  // namespace

extern "C" {

int c4a0_engine_run(c4a0_engine* const* engines, uint32_t n_engines, const c4a0_nn_graph* const* graphs,
                    const uint32_t* n_graphs, void* const* streams, uint64_t max_ticks,
                    uint32_t time_kernels_every, c4a0_run_report* out) {
  if (!graphs) return fail(C4A0_E_INVALID, "bad argument");
  return run_loop(engines, n_engines, graphs, n_graphs, nullptr, streams, max_ticks, time_kernels_every, out);
}

int c4a0_engine_run_net(c4a0_engine* const* engines, uint32_t n_engines, c4a0_net* const* nets, void* const* streams,
                        uint64_t max_ticks, uint32_t time_kernels_every, c4a0_run_report* out) {
  if (!nets) return fail(C4A0_E_INVALID, "bad argument");
  return run_loop(engines, n_engines, nullptr, nullptr, nets, streams, max_ticks, time_kernels_every, out);
}

}  // extern "C"
