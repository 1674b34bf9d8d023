// net.cu — the c4a0 policy/value network as ONE persistent sm_100a kernel (include/c4a0_net.h).
//
// The reference evaluates its network through PyTorch, one library kernel per layer
// (src/c4a0/nn.py:109-130).  In eval mode the network is a chain of dense layers (c4a0_b200/nn.py:
// convolutions on the fixed 6x7 board are constant matrices, BatchNorm folds away), i.e. a handful of
// [rows, K] x [K, N] products with bias + ReLU between them and 7 + 2 outputs at the end.  This file
// runs the whole chain in one launch:
//
//   * work unit = a 128-row x 192-column output tile of one layer.  Tiles are numbered layer by layer,
//     row tile by row tile; CTA b (one per SM, all resident) takes tiles b, b + grid, ...
//   * warp 0 (one lane) is the TMA producer: per 64-wide K step it brings the [128 x 64] activation box
//     and the [192 x 64] weight box (bf16, 128-byte swizzle) into a 5-stage shared-memory ring;
//     warp 1 (one lane) issues tcgen05.mma (M = 128, N = 192, K = 16, f32 accumulators in TMEM, two
//     accumulator buffers so the next tile's MMAs run under this tile's epilogue);
//     warps 2-5 are the epilogue: tcgen05.ld, bias, ReLU, bf16, 16-byte global stores.
//   * no kernel boundary between layers: a tile of layer l+1 needs every column tile of layer l for
//     ITS rows only, so each (layer, row tile) has a counter that epilogues bump and producers wait
//     on.  Row tiles stream through the layers; the tail of one layer overlaps the head of the next.
//   * the two output layers are tiles too (N = 16): their epilogue applies log_softmax / tanh and
//     writes the engine's logits / q buffers (nn.py:116-117).
//   * the number of rows may be read on the device (the engine's row counter), so one launch
//     configuration serves every batch size: no size buckets, no host guess.
//
// Every output element is accumulated over K in a fixed order by the same instruction sequence
// whatever the row's position or the batch size: results are batch-invariant bit for bit.
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/c4a0_net.h"
#include "common.cuh"

namespace {

using c4host::fail;

constexpr uint32_t BM = C4A0_NET_TILE_M, BN = C4A0_NET_TILE_N, BK = C4A0_NET_TILE_K, HEAD_N = C4A0_NET_HEAD_N;
constexpr uint32_t STAGES = 5;
constexpr uint32_t A_BYTES = BM * BK * 2;          // 16 KB
constexpr uint32_t B_BYTES = BN * BK * 2;          // 24 KB
constexpr uint32_t STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t NET_THREADS = 192;              // producer warp, MMA warp, four epilogue warps
constexpr uint32_t TMEM_COLS = 512;                // two accumulators of BN f32 columns (power of two >= 384)
constexpr uint32_t SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /* alignment slack */ + 256 /* barriers */;
constexpr uint32_t MAXL = C4A0_NET_MAX_LAYERS;

struct alignas(64) NetLayer {
  CUtensorMap tmA;  // activations read:  [rows_cap][buffer cols] bf16, box 64 x 128
  CUtensorMap tmW;  // weights:           [n_pad][k_pad] bf16,          box 64 x bn
  const float* bias;
  __nv_bfloat16* out;   // hidden layers: first output element (buffer base + out_col0)
  uint32_t out_stride;  // elements per row of the output buffer
  uint32_t k_blocks, n_tiles, a_col0, bn, kind;
  int32_t dep;
  uint32_t pad[3];
};
struct NetProgram {
  NetLayer layer[MAXL];
};

struct NetArgs {
  NetProgram prog;  // in kernel-parameter space: the TMA unit fetches a descriptor from there without a trip to L2
  uint32_t n_layers, max_mt, rows_cap, rows_fixed;
  const uint32_t *rows_a, *rows_b;
  uint32_t* counters;  // [n_layers][max_mt] finished column tiles of (layer, row tile)
  uint32_t* ticket;    // CTAs that have left the kernel; the last one resets the counters
  int32_t* error;      // set (and the kernel trapped) when a wait did not end: a bug, never a hang
  float *logits, *qp, *qn;
};

// ---- PTX ---------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
constexpr long long WATCHDOG_CYCLES = 6000000000ll;  // ~3 s: a wait that long is a protocol bug
__device__ __noinline__ void watchdog_fire(int32_t* error, int code) {
  *error = code;
  __threadfence_system();
  __trap();
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int32_t* error, int code) {
  if (mbar_try(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try(bar, parity))
    if (clock64() - t0 > WATCHDOG_CYCLES) watchdog_fire(error, code);
}
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Wait until *p >= want (another CTA's release increment), then order what follows — TMA reads of the rows
// that CTA stored, which go through the async proxy — behind it.  The polling loads are relaxed: an acquire
// load per poll would invalidate this SM's L1 every time.
__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t want, int32_t* error);
__device__ __forceinline__ void red_release_add(uint32_t* p, uint32_t v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void prefetch_tensormap(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(m) : "memory");
}
__device__ __noinline__ void watchdog_fire(int32_t* error, int code);
__device__ __forceinline__ void wait_counter(const uint32_t* p, uint32_t want, int32_t* error) {
  if (ld_relaxed(p) < want) {
    const long long t0 = clock64();
    while (ld_relaxed(p) < want) {
      __nanosleep(32);
      if (clock64() - t0 > 6000000000ll) watchdog_fire(error, 1);
    }
  }
  asm volatile("fence.acq_rel.gpu;" ::: "memory");
  asm volatile("fence.proxy.async.global;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// true in exactly one lane of a converged warp
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred)::"memory");
  return pred != 0;
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> f32, one CTA
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// the mbarrier gets one arrival when every MMA issued so far by this thread has completed
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major bf16 tile whose rows are 128 bytes (one swizzle atom
// wide): 8-row groups of 1024 bytes, 128-byte swizzle (what TMA wrote with CU_TENSOR_MAP_SWIZZLE_128B).
//   [0,14) start address >> 4   [16,30) leading byte offset >> 4 (ignored for swizzled K-major, 1)
//   [32,46) stride byte offset >> 4 = 1024 >> 4   [46,48) version = 1 (sm_100)   [61,64) layout 2 = SWIZZLE_128B
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t addr) {
  return (uint64_t)((addr >> 4) & 0x3fffu) | (1ull << 16) | ((uint64_t)(1024u >> 4) << 32) | (1ull << 46) | (2ull << 61);
}
// Instruction descriptor for kind::f16: f32 accumulate (bits 4-5 = 1), A and B bf16 (7-9, 10-12 = 1), both
// K-major (15, 16 = 0), N >> 3 in bits 17-22, M >> 4 in bits 24-28.
__host__ __device__ constexpr uint32_t instr_desc(uint32_t m, uint32_t n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

struct TileId {
  uint32_t layer, m, n;
};

// ---- CTA-pair kernel (cta_group::2) ------------------------------------------------------------
// Measured on B200 with the one-CTA kernel below: every SM pulls ~84 GB/s through its L2 port whatever the
// batch (12.4 TB/s over the chip), and a 128 x 192 tile needs 40 KB per 64-wide K step: the kernel is
// bound by that port, not by the tensor pipe.  Two SMs of a TPC working as one (M = 256: each CTA
// holds its own 128 rows of activations and HALF of the weight tile; the tensor cores read both
// halves) cut the bytes per SM per K step to 16 KB + bn * 64 B for a 128 x bn output per SM.
// Column-tile width is chosen per launch from the row count: 224 (6 tiles per 1,344 columns) for big
// batches, 96 or 32 when a wide tile would leave most SM pairs without work.
constexpr uint32_t NV = 3;
__host__ __device__ constexpr uint32_t variant_bn(uint32_t v) { return v == 0 ? 224u : (v == 1 ? 96u : 32u); }
constexpr uint32_t PAD_N = C4A0_NET_PAD_N;                 // 1344 = lcm(224, 96, 32, 64)
constexpr uint32_t STAGES2 = 6;
constexpr uint32_t NET2_THREADS = 320;                     // producer warp, MMA warp, eight epilogue warps
constexpr uint32_t EPI2_THREADS = 256;
constexpr uint32_t STORE_STAGE_BYTES = 32 * 64;            // per epilogue warp: 32 rows x 32 bf16 columns
constexpr uint32_t B2_BYTES = (224 / 2) * BK * 2;          // 14 KB: half of the widest weight tile
constexpr uint32_t STAGE2_BYTES = A_BYTES + B2_BYTES;      // 30 KB
constexpr uint32_t A32_BYTES = 32 * BK * 2;                // a 32-row activation box (tiny batches)
constexpr uint32_t SMEM2_BYTES = STAGES2 * STAGE2_BYTES + 1024 + 256 + 8 * STORE_STAGE_BYTES;

struct alignas(64) NetLayer2 {
  CUtensorMap tmA128, tmA32;  // activations: box 64 x 128 rows / 64 x 32 rows
  CUtensorMap tmW[NV];        // weights: box 64 x (bn / 2) rows per variant; output layers: tmW[0], 8 rows
  const float* bias;
  __nv_bfloat16* out;
  uint32_t out_stride, k_blocks, n_pad, a_col0, kind;
  int32_t dep;
  uint32_t pad[2];
};
struct NetProgram2 {
  NetLayer2 layer[MAXL];
};
struct NetArgs2 {
  NetProgram2 prog;  // kernel-parameter space (see NetArgs)
  uint32_t n_layers, max_mt, rows_cap, rows_fixed, force_variant;
  uint32_t trace_cta;            // c4a0_net_debug_trace: the CTA whose three roles log (tag, clock64) events ...
  unsigned long long* trace;     // ... into [3][TRACE_N] (nullptr = off)
  const uint32_t *rows_a, *rows_b;
  uint32_t* counters;  // [n_layers][max_mt] finished column tiles of (layer, 128-row tile)
  uint32_t* ticket;
  int32_t* error;
  float *logits, *qp, *qn;
};

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the mbarrier at the same shared-memory offset in CTA `cta` of the cluster.  Relaxed: the barrier
// guards tensor-memory reads, which tcgen05.fence::before_thread_sync orders; a release would also wait for
// this thread's global stores to be acknowledged (~1,500 cycles per tile in the epilogue: measured)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t cta) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(cta)
      : "memory");
}
// TMA load issued by either CTA of a pair; the bytes are credited to the mbarrier of the pair's leader
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int32_t c0, int32_t c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(leader_bar & 0xFEFFFFFFu), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_pair(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// one arrival on the mbarrier at this offset in BOTH CTAs when the MMAs issued so far have completed
__device__ __forceinline__ void umma_commit_pair(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"((uint16_t)3)
               : "memory");
}

constexpr uint32_t TRACE_N = 4096;
struct Tracer {
  unsigned long long* p;
  uint32_t n;
  __device__ __forceinline__ void log(uint32_t tag) {
    if (p != nullptr && n < TRACE_N) p[n++] = ((unsigned long long)clock64() << 8) | (tag & 0xffu);
  }
};

__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NET2_THREADS, 1) k_net2(const __grid_constant__ NetArgs2 A) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t sh_nt[MAXL], sh_kb[MAXL], sh_bn[MAXL], sh_np[MAXL], sh_kind[MAXL], sh_last, sh_variant;
  __shared__ __align__(16) float sh_bias[256];  // the bias slice of the tile the epilogue is working on
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();  // 0 = leader of the pair
  const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar0 = base + STAGES2 * STAGE2_BYTES;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };                        // leader's copy is the live one
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (STAGES2 + s); };           // one per CTA
  auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES2 + a); };       // one per CTA
  auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES2 + 2 + a); };  // leader's copy
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES2 + 4);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES2 * STAGE2_BYTES + 8u * (2 * STAGES2 + 4));

  // ---- prologue that needs nothing from the kernel before this one in the stream: it runs under that
  // kernel's tail when this launch is programmatically dependent (c4a0_net_forward_ex, C4A0_NET_LAUNCH_PDL)
  if (threadIdx.x < A.n_layers) {
    const NetLayer2* L = &A.prog.layer[threadIdx.x];
    sh_kb[threadIdx.x] = L->k_blocks;
    sh_np[threadIdx.x] = L->n_pad;
    sh_kind[threadIdx.x] = L->kind;
    prefetch_tensormap(&L->tmA128);
    prefetch_tensormap(&L->tmA32);
    for (uint32_t v = 0; v < (L->kind == C4A0_NET_HIDDEN ? NV : 1u); v++) prefetch_tensormap(&L->tmW[v]);
  }
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < STAGES2; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 16);  // eight epilogue warps in each CTA of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the MMA warp of each CTA allocates the pair's tensor memory
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // both CTAs' barriers are initialised before anything can arrive on them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  // ---- from here on the kernel reads what its predecessor wrote (the row count, the planes)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  uint32_t rows = A.rows_fixed;
  if (A.rows_a != nullptr) {
    const uint32_t a = *reinterpret_cast<const volatile uint32_t*>(A.rows_a);
    const uint32_t b = *reinterpret_cast<const volatile uint32_t*>(A.rows_b);
    rows = a > b ? a : b;
  }
  if (rows > A.rows_cap) rows = A.rows_cap;
  const uint32_t Mt2 = (rows + 2 * BM - 1) / (2 * BM);  // 256-row blocks: one per CTA pair and column tile
  if (threadIdx.x == 0) {
    // column-tile width of this launch: the variant that moves the fewest bytes through the busiest SM
    uint32_t best = 0;
    unsigned long long best_cost = ~0ull;
    const uint32_t a_eff = rows <= 32u ? A32_BYTES : A_BYTES;
    for (uint32_t v = 0; v < NV; v++) {
      unsigned long long cost = 0;
      for (uint32_t l = 0; l < A.n_layers; l++)
        if (sh_kind[l] == C4A0_NET_HIDDEN) {
          const uint32_t tiles = Mt2 * (sh_np[l] / variant_bn(v));
          cost += (unsigned long long)((tiles + n_clusters - 1) / n_clusters) * sh_kb[l] * (a_eff + variant_bn(v) * 64u);
        }
      if (cost < best_cost) {
        best_cost = cost;
        best = v;
      }
    }
    if (A.force_variant) best = A.force_variant - 1u;
    sh_variant = best;
  }
  __syncthreads();
  const uint32_t variant = sh_variant;
  if (threadIdx.x < A.n_layers) {
    const bool hidden = sh_kind[threadIdx.x] == C4A0_NET_HIDDEN;
    sh_bn[threadIdx.x] = hidden ? variant_bn(variant) : HEAD_N;
    sh_nt[threadIdx.x] = hidden ? sh_np[threadIdx.x] / variant_bn(variant) : 1u;
  }
  __syncthreads();

  uint32_t total = 0;
  for (uint32_t l = 0; l < A.n_layers; l++) total += Mt2 * sh_nt[l];
  auto decode = [&](uint32_t t) {
    TileId id{0, 0, 0};
    for (uint32_t l = 0; l < A.n_layers; l++) {
      const uint32_t c = Mt2 * sh_nt[l];
      if (t < c) {
        id.layer = l;
        id.m = t / sh_nt[l];  // 256-row block
        id.n = t % sh_nt[l];
        break;
      }
      t -= c;
    }
    return id;
  };
  // bytes of the activation box a CTA loads for its 128 rows starting at row0
  auto a_box_bytes = [&](uint32_t row0) { return rows <= row0 ? 0u : (rows - row0 <= 32u ? A32_BYTES : A_BYTES); };

  if (warp == 0) {
    // ===== TMA producer (both CTAs): own 128 rows of activations + own half of the weight tile =====
    // The whole warp runs the loop in lockstep and one elected lane issues: with convergent control flow the
    // descriptors and addresses stay in uniform registers (a divergent single-lane loop costs ~190 cycles
    // per K step in register-to-uniform moves: measured with c4a0_net_debug_trace).
    uint32_t stage = 0, phase = 0;
    Tracer tr{A.trace != nullptr && blockIdx.x == A.trace_cta && lane == 0 ? A.trace : nullptr, 0};
    tr.log(1);
    bool next_ready = false;  // the NEXT tile's dependency was already checked (and acquired) mid-tile
    for (uint32_t t = cluster_id; t < total; t += n_clusters) {
      const TileId id = decode(t);
      const NetLayer2* L = &A.prog.layer[id.layer];
      const uint32_t m = id.m * 2u + rank, row0 = m * BM;
      tr.log(2);
      const uint32_t a_self = a_box_bytes(row0), a_peer = a_box_bytes((id.m * 2u + (rank ^ 1u)) * BM);
      const int32_t dep = L->dep;
      if (dep >= 0 && a_self && !next_ready) {
        if (lane == 0) wait_counter(A.counters + (size_t)dep * A.max_mt + m, sh_nt[dep], A.error);
        __syncwarp();
      }
      next_ready = false;
      tr.log(3);
      const uint32_t kb_n = sh_kb[id.layer], bn = sh_bn[id.layer], half = bn / 2u;
      const bool hidden = sh_kind[id.layer] == C4A0_NET_HIDDEN;
      const CUtensorMap* mapA = a_self == A32_BYTES ? &L->tmA32 : &L->tmA128;
      const CUtensorMap* mapW = &L->tmW[hidden ? variant : 0u];
      const uint32_t bytes = a_self + a_peer + 2u * half * BK * 2u;
      const int32_t a_col0 = (int32_t)L->a_col0, w_row = (int32_t)(id.n * bn + rank * half);
      // look ahead: once the ring is full of this tile's K steps, see whether the next tile's rows are
      // already complete; if so the (expensive) acquire happens here, under the MMAs, not at the tile boundary
      const uint32_t t2 = t + n_clusters;
      const uint32_t peek_at = kb_n > STAGES2 ? STAGES2 : kb_n;
      for (uint32_t kb = 0; kb < kb_n; kb++) {
        mbar_wait(empty_bar(stage), phase ^ 1u, A.error, 2);
        tr.log(4);
        if (elect_one()) {
          if (rank == 0) mbar_expect_tx(full_bar(stage), bytes);
          const uint32_t sa = base + stage * STAGE2_BYTES;
          if (a_self) tma_load_2d_pair(sa, mapA, full_bar(stage), a_col0 + (int32_t)(kb * BK), (int32_t)row0);
          tma_load_2d_pair(sa + A_BYTES, mapW, full_bar(stage), (int32_t)(kb * BK), w_row);
        }
        __syncwarp();
        tr.log(5);
        if (++stage == STAGES2) {
          stage = 0;
          phase ^= 1u;
        }
        if (kb + 1 == peek_at && t2 < total) {
          const TileId nx = decode(t2);
          const int32_t ndep = A.prog.layer[nx.layer].dep;
          const uint32_t nm = nx.m * 2u + rank;
          if (ndep >= 0 && a_box_bytes(nm * BM)) {
            uint32_t ok = 0;
            if (lane == 0) {
              ok = ld_relaxed(A.counters + (size_t)ndep * A.max_mt + nm) >= sh_nt[ndep] ? 1u : 0u;
              if (ok) {
                asm volatile("fence.acq_rel.gpu;" ::: "memory");
                asm volatile("fence.proxy.async.global;" ::: "memory");
              }
            }
            next_ready = __shfl_sync(0xffffffffu, ok, 0) != 0u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer: the leader CTA's warp 1 drives the tensor cores of both SMs (one elected lane issues) =====
    if (rank == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      Tracer tr{A.trace != nullptr && blockIdx.x == A.trace_cta && lane == 0 ? A.trace + TRACE_N : nullptr, 0};
      tr.log(1);
      for (uint32_t t = cluster_id; t < total; t += n_clusters) {
        const TileId id = decode(t);
        const uint32_t kb_n = sh_kb[id.layer], bn = sh_bn[id.layer];
        const uint32_t idesc = instr_desc(2 * BM, bn);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, A.error, 3);  // both CTAs' epilogues have drained this accumulator
        tr.log(2);
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * 224u;
        for (uint32_t kb = 0; kb < kb_n; kb++) {
          mbar_wait(full_bar(stage), phase, A.error, 4);
          tr.log(4);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE2_BYTES;
          const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + A_BYTES);
          if (elect_one()) {
#pragma unroll
            for (uint32_t kk = 0; kk < BK / 16; kk++)
              umma_bf16_pair(tmem_d, adesc + 2u * kk, bdesc + 2u * kk, idesc, (kb | kk) != 0u ? 1u : 0u);
            umma_commit_pair(empty_bar(stage));  // both producers may refill the stage
            if (kb + 1 == kb_n) umma_commit_pair(tfull_bar(acc));  // both epilogues may read their 128 rows
          }
          __syncwarp();
          tr.log(5);
          if (++stage == STAGES2) {
            stage = 0;
            phase ^= 1u;
          }
        }
        acc ^= 1u;
        if (acc == 0u) acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue (both CTAs): this CTA's 128 rows of the tile =====
    // warps 2..9: warp w may touch TMEM lanes [32 (w % 4), +32); the two warps of a lane quarter take
    // alternate 32-column chunks of the tile
    const uint32_t q = warp & 3u;
    const uint32_t epi_tid = threadIdx.x - 64u;
    const uint32_t chunk0 = (warp - 2u) >> 2;  // 0 or 1
    // per-warp staging buffer for coalesced stores: 32 rows x 64 bytes, 16-byte pieces XOR-swizzled
    uint8_t* stage_gen = smem_gen + STAGES2 * STAGE2_BYTES + 256 + (warp - 2u) * STORE_STAGE_BYTES;
    uint32_t acc = 0, acc_phase = 0;
    Tracer tr{A.trace != nullptr && blockIdx.x == A.trace_cta && epi_tid == 0 ? A.trace + 2 * TRACE_N : nullptr, 0};
    tr.log(1);
    for (uint32_t t = cluster_id; t < total; t += n_clusters) {
      const TileId id = decode(t);
      const NetLayer2* L = &A.prog.layer[id.layer];
      const uint32_t kind = sh_kind[id.layer], bn = sh_bn[id.layer];
      tr.log(2);
      const uint32_t m = id.m * 2u + rank;
      {  // stage the tile's biases (the closing barrier of the previous tile protects the buffer)
        const float* src = L->bias + (kind == C4A0_NET_HIDDEN ? id.n * bn : 0u);
        if (epi_tid < bn) sh_bias[epi_tid] = __ldg(src + epi_tid);
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
      mbar_wait(tfull_bar(acc), acc_phase, A.error, 5);
      tr.log(6);
      tc_fence_after();
      const uint32_t row = m * BM + q * 32u + lane;
      const bool live = row < rows;  // rows past the batch hold whatever the stage buffers held: never stored
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * 224u;
      const bool warp_live = m * BM + q * 32u < rows;  // warp-uniform: any of this warp's 32 rows in the batch?
      if (kind == C4A0_NET_HIDDEN) {
        const uint32_t chunks = bn / 32u;
        const uint32_t row_base = m * BM + q * 32u;
        __nv_bfloat16* out = L->out + (size_t)row_base * L->out_stride + id.n * bn;
        const uint32_t out_stride = L->out_stride;
        // bias + ReLU + bf16 of 32 accumulator columns of this lane's row; the warp's 32 x 32 block goes
        // through shared memory so that global stores are whole 32-byte sectors: 4 lanes per row, 8 rows
        // per instruction (a lane storing its own row directly writes 32 scattered half sectors)
        auto emit = [&](const uint32_t (&v)[32], uint32_t c) {
          uint4* st = reinterpret_cast<uint4*>(stage_gen + lane * 64u);
          const uint32_t f = (lane >> 1) & 3u;
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 b2 = *reinterpret_cast<const float2*>(&sh_bias[c * 32u + g * 8 + j * 2]);
              const float x0 = fmaxf(__uint_as_float(v[g * 8 + j * 2]) + b2.x, 0.0f);
              const float x1 = fmaxf(__uint_as_float(v[g * 8 + j * 2 + 1]) + b2.y, 0.0f);
              __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            st[(uint32_t)g ^ f] = make_uint4(w[0], w[1], w[2], w[3]);
          }
          __syncwarp();
          tr.log(10);
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const uint32_t r = 8u * i + (lane >> 2), piece = lane & 3u;
            const uint4 d = *reinterpret_cast<const uint4*>(stage_gen + r * 64u + ((piece ^ ((r >> 1) & 3u)) << 4));
            if (row_base + r < rows)
              *reinterpret_cast<uint4*>(out + (size_t)r * out_stride + c * 32u + piece * 8u) = d;
          }
          __syncwarp();
          tr.log(11);
        };
        if (warp_live) {  // two register buffers: the next chunk's TMEM load runs under this chunk's math and stores
          uint32_t va[32], vb[32];
          if (chunk0 < chunks) tmem_ld32(taddr + chunk0 * 32u, va);
#pragma unroll 1
          for (uint32_t c = chunk0; c < chunks; c += 4) {
            tmem_ld_wait();
            tr.log(9);
            if (c + 2 < chunks) tmem_ld32(taddr + (c + 2) * 32u, vb);
            emit(va, c);
            if (c + 2 < chunks) {
              tmem_ld_wait();
              tr.log(9);
              if (c + 4 < chunks) tmem_ld32(taddr + (c + 4) * 32u, va);
              emit(vb, c + 2);
            }
          }
        }
      } else {
        uint32_t v[16];
        if (warp_live && chunk0 == 0u) {
          tmem_ld16(taddr, v);
          tmem_ld_wait();
        }
        if (live && chunk0 == 0u) {
          if (kind == C4A0_NET_POLICY) {  // nn.py:84-86: LogSoftmax over the seven columns
            float x[7], mx = -3.402823466e38f;
#pragma unroll
            for (int k = 0; k < 7; k++) {
              x[k] = __uint_as_float(v[k]) + sh_bias[k];
              mx = fmaxf(mx, x[k]);
            }
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 7; k++) s += expf(x[k] - mx);
            const float lse = mx + logf(s);
#pragma unroll
            for (int k = 0; k < 7; k++) A.logits[(size_t)row * 7 + k] = x[k] - lse;
          } else {  // nn.py:98-100: Tanh of the two values
            A.qp[row] = tanhf(__uint_as_float(v[0]) + sh_bias[0]);
            A.qn[row] = tanhf(__uint_as_float(v[1]) + sh_bias[1]);
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(tempty_bar(acc), 0u);  // the leader's barrier counts both CTAs
      tr.log(7);
      acc ^= 1u;
      if (acc == 0u) acc_phase ^= 1u;
      // publish the tile: every thread orders its stores before later async-proxy (TMA) reads, the 256
      // epilogue threads meet, one release increment makes the rows visible to the waiting producers
      asm volatile("fence.proxy.async.global;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (epi_tid == 0) red_release_add(A.counters + (size_t)id.layer * A.max_mt + m, 1u);
      tr.log(8);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  cluster_sync();  // neither CTA frees tensor memory (or exits) while its peer may still touch it
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  if (threadIdx.x == 0) {
    __threadfence();
    sh_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last) {
    __threadfence();
    const uint32_t n = A.n_layers * A.max_mt;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) A.counters[i] = 0u;
    if (threadIdx.x == 0) *A.ticket = 0u;
  }
}

// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(NET_THREADS, 1) k_net(const __grid_constant__ NetArgs A) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ uint32_t sh_nt[MAXL], sh_kb[MAXL], sh_bn[MAXL], sh_last;
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // swizzle-128B tiles need 1024-byte alignment
  const uint32_t bar0 = base + STAGES * STAGE_BYTES;
  auto full_bar = [&](uint32_t s) { return bar0 + 8u * s; };
  auto empty_bar = [&](uint32_t s) { return bar0 + 8u * (STAGES + s); };
  auto tfull_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES + a); };
  auto tempty_bar = [&](uint32_t a) { return bar0 + 8u * (2 * STAGES + 2 + a); };
  const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 4);
  uint8_t* smem_gen = smem_raw + (base - smem_u32(smem_raw));
  volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8u * (2 * STAGES + 4));

  uint32_t rows = A.rows_fixed;
  if (A.rows_a != nullptr) {
    const uint32_t a = *reinterpret_cast<const volatile uint32_t*>(A.rows_a);
    const uint32_t b = *reinterpret_cast<const volatile uint32_t*>(A.rows_b);
    rows = a > b ? a : b;
  }
  if (rows > A.rows_cap) rows = A.rows_cap;
  const uint32_t Mt = (rows + BM - 1) / BM;

  if (threadIdx.x < A.n_layers) {
    sh_nt[threadIdx.x] = A.prog.layer[threadIdx.x].n_tiles;
    sh_kb[threadIdx.x] = A.prog.layer[threadIdx.x].k_blocks;
    sh_bn[threadIdx.x] = A.prog.layer[threadIdx.x].bn;
  }
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < STAGES; s++) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; a++) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4);  // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // the MMA warp owns the tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(TMEM_COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_gen;

  uint32_t total = 0;
  for (uint32_t l = 0; l < A.n_layers; l++) total += Mt * sh_nt[l];
  auto decode = [&](uint32_t t) {
    TileId id{0, 0, 0};
    for (uint32_t l = 0; l < A.n_layers; l++) {
      const uint32_t c = Mt * sh_nt[l];
      if (t < c) {
        id.layer = l;
        id.m = t / sh_nt[l];
        id.n = t % sh_nt[l];
        break;
      }
      t -= c;
    }
    return id;
  };

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0;
      for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        const TileId id = decode(t);
        const NetLayer* L = &A.prog.layer[id.layer];
        const int32_t dep = L->dep;
        // the rows this tile reads are complete once every column tile of `dep` is stored
        if (dep >= 0) wait_counter(A.counters + (size_t)dep * A.max_mt + id.m, sh_nt[dep], A.error);
        const uint32_t kb_n = sh_kb[id.layer], bn = sh_bn[id.layer];
        const uint32_t bytes = A_BYTES + bn * BK * 2;
        const int32_t a_col0 = (int32_t)L->a_col0;
        for (uint32_t kb = 0; kb < kb_n; kb++) {
          mbar_wait(empty_bar(stage), phase ^ 1u, A.error, 2);
          mbar_expect_tx(full_bar(stage), bytes);
          const uint32_t sa = base + stage * STAGE_BYTES;
          tma_load_2d(sa, &L->tmA, full_bar(stage), a_col0 + (int32_t)(kb * BK), (int32_t)(id.m * BM));
          tma_load_2d(sa + A_BYTES, &L->tmW, full_bar(stage), (int32_t)(kb * BK), (int32_t)(id.n * bn));
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer =====
    if (lane == 0) {
      uint32_t stage = 0, phase = 0, acc = 0, acc_phase = 0;
      for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
        const TileId id = decode(t);
        const uint32_t kb_n = sh_kb[id.layer], bn = sh_bn[id.layer];
        const uint32_t idesc = instr_desc(BM, bn);
        mbar_wait(tempty_bar(acc), acc_phase ^ 1u, A.error, 3);  // the epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BN;
        for (uint32_t kb = 0; kb < kb_n; kb++) {
          mbar_wait(full_bar(stage), phase, A.error, 4);
          tc_fence_after();
          const uint32_t sa = base + stage * STAGE_BYTES;
          const uint64_t adesc = smem_desc_sw128(sa), bdesc = smem_desc_sw128(sa + A_BYTES);
#pragma unroll
          for (uint32_t k = 0; k < BK / 16; k++)  // 32 bytes along K per instruction: start address + 2 (x16 bytes)
            umma_bf16(tmem_d, adesc + 2u * k, bdesc + 2u * k, idesc, (kb | k) != 0u ? 1u : 0u);
          umma_commit(empty_bar(stage));  // frees the stage when these MMAs have read it
          if (++stage == STAGES) {
            stage = 0;
            phase ^= 1u;
          }
        }
        umma_commit(tfull_bar(acc));  // accumulator complete
        acc ^= 1u;
        if (acc == 0u) acc_phase ^= 1u;
      }
    }
  } else {
    // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes [32 (warp % 4), +32) =====
    const uint32_t q = warp & 3u;
    const uint32_t epi_tid = threadIdx.x - 64u;
    uint32_t acc = 0, acc_phase = 0;
    for (uint32_t t = blockIdx.x; t < total; t += gridDim.x) {
      const TileId id = decode(t);
      const NetLayer* L = &A.prog.layer[id.layer];
      const uint32_t kind = L->kind;
      mbar_wait(tfull_bar(acc), acc_phase, A.error, 5);
      tc_fence_after();
      const uint32_t row = id.m * BM + q * 32u + lane;
      const uint32_t taddr = tmem_base + ((q * 32u) << 16) + acc * BN;
      if (kind == C4A0_NET_HIDDEN) {
        const float* bias = L->bias + id.n * BN;
        __nv_bfloat16* out = L->out + (size_t)row * L->out_stride + id.n * BN;
#pragma unroll 1
        for (uint32_t c = 0; c < BN / 32; c++) {
          uint32_t v[32];
          tmem_ld32(taddr + c * 32u, v);
          tmem_ld_wait();
          uint4* dst = reinterpret_cast<uint4*>(out + c * 32u);
#pragma unroll
          for (int g = 0; g < 4; g++) {
            uint32_t w[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
              const float2 b2 = __ldg(reinterpret_cast<const float2*>(bias + c * 32u + g * 8 + j * 2));
              const float x0 = fmaxf(__uint_as_float(v[g * 8 + j * 2]) + b2.x, 0.0f);
              const float x1 = fmaxf(__uint_as_float(v[g * 8 + j * 2 + 1]) + b2.y, 0.0f);
              __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
              w[j] = *reinterpret_cast<uint32_t*>(&h);
            }
            dst[g] = make_uint4(w[0], w[1], w[2], w[3]);
          }
        }
      } else {
        uint32_t v[16];
        tmem_ld16(taddr, v);
        tmem_ld_wait();
        const float* bias = L->bias;
        if (row < rows) {
          if (kind == C4A0_NET_POLICY) {  // nn.py:84-86: LogSoftmax over the seven columns
            float x[7], mx = -3.402823466e38f;
#pragma unroll
            for (int k = 0; k < 7; k++) {
              x[k] = __uint_as_float(v[k]) + __ldg(bias + k);
              mx = fmaxf(mx, x[k]);
            }
            float s = 0.0f;
#pragma unroll
            for (int k = 0; k < 7; k++) s += expf(x[k] - mx);
            const float lse = mx + logf(s);
#pragma unroll
            for (int k = 0; k < 7; k++) A.logits[(size_t)row * 7 + k] = x[k] - lse;
          } else {  // nn.py:98-100: Tanh of the two values
            A.qp[row] = tanhf(__uint_as_float(v[0]) + __ldg(bias + 0));
            A.qn[row] = tanhf(__uint_as_float(v[1]) + __ldg(bias + 1));
          }
        }
      }
      // the accumulator can be overwritten
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(tempty_bar(acc));
      acc ^= 1u;
      if (acc == 0u) acc_phase ^= 1u;
      asm volatile("fence.proxy.async.global;" ::: "memory");
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (epi_tid == 0) red_release_add(A.counters + (size_t)id.layer * A.max_mt + id.m, 1u);
    }
  }

  // ---- teardown ----
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(TMEM_COLS) : "memory");
  }
  if (threadIdx.x == 0) {
    __threadfence();
    sh_last = atomicAdd(A.ticket, 1u) == gridDim.x - 1 ? 1u : 0u;
  }
  __syncthreads();
  if (sh_last) {  // every CTA is done: clear the counters for the next launch
    __threadfence();
    const uint32_t n = A.n_layers * A.max_mt;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) A.counters[i] = 0u;
    if (threadIdx.x == 0) *A.ticket = 0u;
  }
}

// ---- host --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

// [rows][cols] bf16 row-major, box = 64 columns x box_rows rows, 128-byte swizzle
int make_map(CUtensorMap* m, const void* base, uint64_t rows, uint64_t cols, uint32_t box_rows) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(C4A0_E_CUDA, "cuTensorMapEncodeTiled is not available from this driver");
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {cols * 2};
  cuuint32_t box[2] = {BK, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(C4A0_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return 0;
}

}  // namespace

struct c4a0_net {
  c4a0_net_spec spec;
  std::vector<void*> allocs;
  size_t bytes = 0;
  void* buffers[C4A0_NET_MAX_BUFFERS] = {};
  uint32_t rows_cap = 0, grid = 0;
  bool pair = true;  // CTA-pair kernel (k_net2); the one-CTA kernel (k_net) is kept for comparison
  NetArgs args{};
  NetArgs2 args2{};
  int32_t* h_error = nullptr;  // mapped pinned
  cudaEvent_t ev[2] = {nullptr, nullptr};
};

extern "C" {

void c4a0_net_destroy(c4a0_net* n) {
  if (!n) return;
  cudaSetDevice(n->spec.device);
  cudaDeviceSynchronize();
  for (void* p : n->allocs) cudaFree(p);
  if (n->h_error) cudaFreeHost(n->h_error);
  for (auto e : n->ev)
    if (e) cudaEventDestroy(e);
  delete n;
}

int c4a0_net_create(const c4a0_net_spec* spec, c4a0_net** out) {
  if (!spec || !out) return fail(C4A0_E_INVALID, "null argument");
  *out = nullptr;
  if (spec->n_layers == 0 || spec->n_layers > MAXL || spec->n_buffers == 0 || spec->n_buffers > C4A0_NET_MAX_BUFFERS)
    return fail(C4A0_E_INVALID, "bad layer / buffer count");
  if (spec->max_rows == 0) return fail(C4A0_E_INVALID, "max_rows must be >= 1");
  for (uint32_t b = 0; b < spec->n_buffers; b++)
    if (spec->buffer_cols[b] == 0 || spec->buffer_cols[b] % BK) return fail(C4A0_E_INVALID, "buffer_cols must be multiples of 64");
  if (spec->planes_buffer >= spec->n_buffers || spec->planes_col0 % 8 || spec->planes_col0 + 88 > spec->buffer_cols[spec->planes_buffer])
    return fail(C4A0_E_INVALID, "bad planes location");
  bool has_policy = false, has_value = false;
  for (uint32_t l = 0; l < spec->n_layers; l++) {
    const c4a0_net_layer& L = spec->layers[l];
    const bool head = L.kind != C4A0_NET_HIDDEN;
    if (L.kind > C4A0_NET_VALUE) return fail(C4A0_E_INVALID, "layer %u: bad kind", l);
    if (!L.weight_dev || !L.bias_dev) return fail(C4A0_E_INVALID, "layer %u: null weights", l);
    if (((uintptr_t)L.weight_dev & 15u) || ((uintptr_t)L.bias_dev & 15u)) return fail(C4A0_E_INVALID, "layer %u: weights must be 16-byte aligned", l);
    if (L.k_pad == 0 || L.k_pad % BK) return fail(C4A0_E_INVALID, "layer %u: k_pad must be a multiple of 64", l);
    if (head ? L.n_pad != HEAD_N : (L.n_pad == 0 || L.n_pad % PAD_N)) return fail(C4A0_E_INVALID, "layer %u: bad n_pad", l);
    if (L.in_buffer >= spec->n_buffers || L.in_col0 % BK || L.in_col0 + L.k_pad > spec->buffer_cols[L.in_buffer])
      return fail(C4A0_E_INVALID, "layer %u: input columns out of range", l);
    if (!head && (L.out_buffer >= spec->n_buffers || L.out_col0 % 8 || L.out_col0 + L.n_pad > spec->buffer_cols[L.out_buffer]))
      return fail(C4A0_E_INVALID, "layer %u: output columns out of range", l);
    if (L.dep >= (int32_t)l || L.dep < -1) return fail(C4A0_E_INVALID, "layer %u: dep must name an earlier layer or be -1", l);
    if (L.dep >= 0 && spec->layers[L.dep].kind != C4A0_NET_HIDDEN) return fail(C4A0_E_INVALID, "layer %u: dep is an output layer", l);
    has_policy |= L.kind == C4A0_NET_POLICY;
    has_value |= L.kind == C4A0_NET_VALUE;
  }
  if (!has_policy || !has_value) return fail(C4A0_E_INVALID, "the program needs a policy and a value output layer");
  int r = c4host::no_gpu_error();
  if (r) return r;
  CK(cudaSetDevice(spec->device));
  int cc_major = 0, sms = 0;
  CK(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, spec->device));
  CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, spec->device));
  if (cc_major != 10) return fail(C4A0_E_CUDA, "the network kernel is written for sm_100a (tcgen05/TMEM/TMA); this device is sm_%d", cc_major);
  c4a0_net* n = new c4a0_net();
  n->spec = *spec;
  n->rows_cap = (spec->max_rows + 2 * BM - 1) / (2 * BM) * (2 * BM);
  n->grid = (uint32_t)sms & ~1u;
  if (const char* env = getenv("C4A0_NET_ONE_CTA")) n->pair = !(env[0] == '1');
  if (!n->pair)
    for (uint32_t l = 0; l < spec->n_layers; l++)
      if (spec->layers[l].kind == C4A0_NET_HIDDEN && spec->layers[l].n_pad % BN) {
        delete n;
        return fail(C4A0_E_INVALID, "the one-CTA kernel needs hidden widths that are multiples of %u", BN);
      }
  auto dalloc = [&](void** p, size_t b) -> int {
    cudaError_t e = cudaMalloc(p, b ? b : 1);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(e == cudaErrorMemoryAllocation ? C4A0_E_NOMEM : C4A0_E_CUDA, "cudaMalloc(%zu bytes) failed: %s", b, cudaGetErrorString(e));
    }
    n->allocs.push_back(*p);
    n->bytes += b;
    return 0;
  };
#define NA(expr)              \
  do {                        \
    int _r = (expr);          \
    if (_r) {                 \
      c4a0_net_destroy(n);    \
      return _r;              \
    }                         \
  } while (0)
#define NCK(call)                                                                                         \
  do {                                                                                                    \
    cudaError_t _e = (call);                                                                              \
    if (_e != cudaSuccess) {                                                                              \
      c4a0_net_destroy(n);                                                                                \
      return fail(C4A0_E_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(_e), __FILE__, __LINE__); \
    }                                                                                                     \
  } while (0)
  for (uint32_t b = 0; b < spec->n_buffers; b++) {
    const size_t bytes = (size_t)n->rows_cap * spec->buffer_cols[b] * 2;
    NA(dalloc(&n->buffers[b], bytes));
    NCK(cudaMemset(n->buffers[b], 0, bytes));
  }
  const uint32_t max_mt = n->rows_cap / BM;
  void* p = nullptr;
  NA(dalloc(&p, (size_t)spec->n_layers * max_mt * 4));
  uint32_t* counters = reinterpret_cast<uint32_t*>(p);
  NCK(cudaMemset(counters, 0, (size_t)spec->n_layers * max_mt * 4));
  NA(dalloc(&p, 4));
  uint32_t* ticket = reinterpret_cast<uint32_t*>(p);
  NCK(cudaMemset(ticket, 0, 4));
  if (n->pair) {
    NetArgs2& A = n->args2;
    memset(&A.prog, 0, sizeof(A.prog));
    for (uint32_t l = 0; l < spec->n_layers; l++) {
      const c4a0_net_layer& L = spec->layers[l];
      NetLayer2& D = A.prog.layer[l];
      const bool head = L.kind != C4A0_NET_HIDDEN;
      NA(make_map(&D.tmA128, n->buffers[L.in_buffer], n->rows_cap, spec->buffer_cols[L.in_buffer], BM));
      NA(make_map(&D.tmA32, n->buffers[L.in_buffer], n->rows_cap, spec->buffer_cols[L.in_buffer], 32));
      for (uint32_t v = 0; v < (head ? 1u : NV); v++)
        NA(make_map(&D.tmW[v], L.weight_dev, L.n_pad, L.k_pad, head ? HEAD_N / 2 : variant_bn(v) / 2));
      D.bias = L.bias_dev;
      D.out = head ? nullptr : reinterpret_cast<__nv_bfloat16*>(n->buffers[L.out_buffer]) + L.out_col0;
      D.out_stride = head ? 0 : spec->buffer_cols[L.out_buffer];
      D.k_blocks = L.k_pad / BK;
      D.n_pad = L.n_pad;
      D.a_col0 = L.in_col0;
      D.kind = L.kind;
      D.dep = L.dep;
    }
    A.n_layers = spec->n_layers;
    A.max_mt = max_mt;
    A.rows_cap = spec->max_rows;
    A.counters = counters;
    A.ticket = ticket;
    if (const char* env = getenv("C4A0_NET_VARIANT")) A.force_variant = (uint32_t)atoi(env) <= NV ? (uint32_t)atoi(env) : 0u;
    NCK(cudaFuncSetAttribute(k_net2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM2_BYTES));
  } else {
    NetArgs& A = n->args;
    memset(&A.prog, 0, sizeof(A.prog));
    for (uint32_t l = 0; l < spec->n_layers; l++) {
      const c4a0_net_layer& L = spec->layers[l];
      NetLayer& D = A.prog.layer[l];
      const bool head = L.kind != C4A0_NET_HIDDEN;
      D.bn = head ? HEAD_N : BN;
      NA(make_map(&D.tmA, n->buffers[L.in_buffer], n->rows_cap, spec->buffer_cols[L.in_buffer], BM));
      NA(make_map(&D.tmW, L.weight_dev, L.n_pad, L.k_pad, D.bn));
      D.bias = L.bias_dev;
      D.out = head ? nullptr : reinterpret_cast<__nv_bfloat16*>(n->buffers[L.out_buffer]) + L.out_col0;
      D.out_stride = head ? 0 : spec->buffer_cols[L.out_buffer];
      D.k_blocks = L.k_pad / BK;
      D.n_tiles = L.n_pad / D.bn;
      D.a_col0 = L.in_col0;
      D.kind = L.kind;
      D.dep = L.dep;
    }
    A.n_layers = spec->n_layers;
    A.max_mt = max_mt;
    A.rows_cap = spec->max_rows;
    A.counters = counters;
    A.ticket = ticket;
    NCK(cudaFuncSetAttribute(k_net, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
  }
  NCK(cudaHostAlloc((void**)&n->h_error, 4, cudaHostAllocMapped));
  *n->h_error = 0;
  {
    int32_t* derr = nullptr;
    NCK(cudaHostGetDevicePointer((void**)&derr, n->h_error, 0));
    n->args.error = derr;
    n->args2.error = derr;
  }
  NCK(cudaDeviceSynchronize());
#undef NA
#undef NCK
  *out = n;
  return 0;
}

size_t c4a0_net_device_bytes(const c4a0_net* n) { return n ? n->bytes : 0; }

int c4a0_net_buffer(c4a0_net* n, uint32_t buffer, void** base_dev, uint32_t* cols, uint32_t* rows) {
  if (!n || buffer >= n->spec.n_buffers) return fail(C4A0_E_INVALID, "bad buffer index");
  if (base_dev) *base_dev = n->buffers[buffer];
  if (cols) *cols = n->spec.buffer_cols[buffer];
  if (rows) *rows = n->rows_cap;
  return 0;
}

int c4a0_net_bind_outputs(c4a0_net* n, float* logits, float* qp, float* qn) {
  if (!n || !logits || !qp || !qn) return fail(C4A0_E_INVALID, "null argument");
  n->args.logits = n->args2.logits = logits;
  n->args.qp = n->args2.qp = qp;
  n->args.qn = n->args2.qn = qn;
  return 0;
}

int c4a0_net_bind_row_count(c4a0_net* n, const uint32_t* a, const uint32_t* b) {
  if (!n || ((a == nullptr) != (b == nullptr))) return fail(C4A0_E_INVALID, "pass both counters or neither");
  n->args.rows_a = n->args2.rows_a = a;
  n->args.rows_b = n->args2.rows_b = b;
  return 0;
}

int c4a0_net_forward(c4a0_net* n, uint32_t rows, void* stream) { return c4a0_net_forward_ex(n, rows, stream, 0u); }

int c4a0_net_forward_ex(c4a0_net* n, uint32_t rows, void* stream, uint32_t flags) {
  if (!n) return fail(C4A0_E_INVALID, "null net");
  if (!n->args.logits) return fail(C4A0_E_INVALID, "bind_outputs() must precede forward()");
  if (rows > n->spec.max_rows) return fail(C4A0_E_INVALID, "%u rows exceed max_rows=%u", rows, n->spec.max_rows);
  if (*n->h_error) return fail(C4A0_E_ENGINE, "the network kernel reported a stuck wait (code %d)", *n->h_error);
  if (n->pair) {
    n->args2.rows_fixed = rows;
    if (flags & C4A0_NET_LAUNCH_PDL) {
      // programmatic dependent launch: the CTAs may become resident, and run their prologue, while the kernel
      // before this one in the stream is still draining; griddepcontrol.wait in the kernel orders the rest
      cudaLaunchConfig_t cfg = {};
      cfg.gridDim = dim3(n->grid);
      cfg.blockDim = dim3(NET2_THREADS);
      cfg.dynamicSmemBytes = SMEM2_BYTES;
      cfg.stream = (cudaStream_t)stream;
      cudaLaunchAttribute at[1];
      at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      at[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = at;
      cfg.numAttrs = 1;
      CK(cudaLaunchKernelEx(&cfg, k_net2, n->args2));
    } else {
      k_net2<<<n->grid, NET2_THREADS, SMEM2_BYTES, (cudaStream_t)stream>>>(n->args2);  // clusters of two (__cluster_dims__)
    }
  } else {
    n->args.rows_fixed = rows;
    k_net<<<n->grid, NET_THREADS, SMEM_BYTES, (cudaStream_t)stream>>>(n->args);
  }
  CK(cudaGetLastError());
  return 0;
}

int c4a0_net_debug_trace(c4a0_net* n, uint32_t rows, uint32_t cta, void* stream, uint64_t* out, size_t n_out) {
  if (!n || !out) return fail(C4A0_E_INVALID, "null argument");
  if (!n->pair) return fail(C4A0_E_INVALID, "tracing is implemented for the CTA-pair kernel");
  if (n_out < 3 * TRACE_N) return fail(C4A0_E_INVALID, "the trace buffer needs %u entries", 3 * TRACE_N);
  cudaStream_t s = (cudaStream_t)stream;
  unsigned long long* d = nullptr;
  CK(cudaMalloc((void**)&d, 3 * TRACE_N * 8));
  CK(cudaMemsetAsync(d, 0, 3 * TRACE_N * 8, s));
  n->args2.trace = d;
  n->args2.trace_cta = cta;
  int r = c4a0_net_forward(n, rows, stream);
  n->args2.trace = nullptr;
  if (!r) {
    cudaError_t e = cudaMemcpyAsync(out, d, 3 * TRACE_N * 8, cudaMemcpyDeviceToHost, s);
    if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    if (e != cudaSuccess) r = fail(C4A0_E_CUDA, "trace copy failed: %s", cudaGetErrorString(e));
  }
  cudaFree(d);
  return r;
}

int c4a0_net_forward_timed(c4a0_net* n, uint32_t rows, void* stream, float* ms) {
  if (!n) return fail(C4A0_E_INVALID, "null net");
  cudaStream_t s = (cudaStream_t)stream;
  if (!n->ev[0])
    for (int i = 0; i < 2; i++) CK(cudaEventCreate(&n->ev[i]));
  CK(cudaEventRecord(n->ev[0], s));
  int r = c4a0_net_forward(n, rows, stream);
  if (r) return r;
  CK(cudaEventRecord(n->ev[1], s));
  CK(cudaStreamSynchronize(s));
  float t = 0;
  CK(cudaEventElapsedTime(&t, n->ev[0], n->ev[1]));
  if (ms) *ms = t;
  if (*n->h_error) return fail(C4A0_E_ENGINE, "the network kernel reported a stuck wait (code %d)", *n->h_error);
  return 0;
}

}  // extern "C"
