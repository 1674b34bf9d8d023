/*
 * c4a0_net.h — C-ABI of the hand-written sm_100a evaluator for the c4a0 policy/value network
 * (part of libc4a0_engine.so).
 *
 * What it replaces: in the reference every tick of self-play calls back into Python,
 * `ConnectFourNet.forward_numpy` (src/c4a0/nn.py:119-130) -> `forward` (nn.py:109-117): cuDNN
 * convolutions, cuBLAS linears, BatchNorm, LogSoftmax/Tanh, one library kernel per layer, with a
 * host<->device copy either side (rust/src/pybridge.rs:161-199).  In eval mode that function is a
 * chain of dense layers (convolutions on the fixed 6x7 board are constant matrices, BatchNorm folds
 * into the layer before it; see c4a0_b200/nn.py FoldedNet/FusedNet), and this library evaluates the
 * whole chain in ONE persistent kernel: TMA-staged bf16 tiles, tcgen05.mma with TMEM accumulators,
 * bias + ReLU + bf16 rounding in the epilogue, the two output layers with log_softmax / tanh fused,
 * and per-row-tile dependency counters instead of kernel boundaries between layers.
 *
 * The caller describes the chain as a short program of dense layers over a few activation buffers
 * (`c4a0_net_spec`); weights stay in caller-owned device memory (so a new generation's weights are
 * loaded by overwriting them in place).  Every output row depends only on its own input row and is
 * accumulated in a fixed order, so results are bit-identical whatever the batch size or the row's
 * position in the batch.
 *
 * Conventions as in c4a0_engine.h: plain C types, 0 / negative C4A0_E_* return codes with
 * c4a0_last_error(), `*_dev` = device pointers, `stream` = cudaStream_t as void*.
 */
#ifndef C4A0_NET_H
#define C4A0_NET_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define C4A0_NET_MAX_LAYERS 12
#define C4A0_NET_MAX_BUFFERS 8
#define C4A0_NET_TILE_M 128   /* rows per tile */
#define C4A0_NET_TILE_N 192   /* output columns per tile of a hidden layer (one-CTA kernel) */
#define C4A0_NET_PAD_N 1344   /* hidden widths are padded to a multiple of this: a common multiple of the CTA-pair
                                 kernel's column tiles (224, 96, 32) and of the 64-wide K step */
#define C4A0_NET_TILE_K 64    /* input columns per pipeline stage */
#define C4A0_NET_HEAD_N 16    /* padded width of an output layer (7 logits / 2 values) */

enum {
  C4A0_NET_HIDDEN = 0, /* out = bf16(relu(in @ W^T + b))                           nn.py:75-100 (Linear+BN+ReLU) */
  C4A0_NET_POLICY = 1, /* logits = log_softmax(in @ W^T + b) over 7 columns        nn.py:84-86, 116 */
  C4A0_NET_VALUE = 2   /* (q_penalty, q_no_penalty) = tanh(in @ W^T + b), 2 cols   nn.py:98-100, 117 */
};

typedef struct {
  const void* weight_dev;  /* bf16 [n_pad][k_pad] row-major (k contiguous), zero padded */
  const float* bias_dev;   /* f32 [n_pad], zero padded */
  uint32_t n_pad;          /* hidden: a multiple of C4A0_NET_PAD_N; output layers: C4A0_NET_HEAD_N */
  uint32_t k_pad;          /* a multiple of C4A0_NET_TILE_K */
  uint32_t in_buffer;      /* activation buffer read ... */
  uint32_t in_col0;        /* ... from this column on (a multiple of 64), k_pad columns */
  uint32_t out_buffer;     /* hidden: activation buffer written ... */
  uint32_t out_col0;       /* ... from this column on (a multiple of 8), n_pad columns */
  int32_t dep;             /* index (< this layer's) of the layer that produces the input, -1 = the input planes */
  uint32_t kind;           /* C4A0_NET_* */
} c4a0_net_layer;

typedef struct {
  int32_t device;
  uint32_t max_rows;                              /* capacity of a forward pass */
  uint32_t n_buffers;
  uint32_t buffer_cols[C4A0_NET_MAX_BUFFERS];     /* bf16 columns per row of each activation buffer (multiples of 64) */
  uint32_t planes_buffer, planes_col0;            /* where the caller (the self-play engine) writes the 84 input planes
                                                     of a row: bf16, row stride buffer_cols[planes_buffer] */
  uint32_t n_layers;
  c4a0_net_layer layers[C4A0_NET_MAX_LAYERS];
} c4a0_net_spec;

typedef struct c4a0_net c4a0_net;

/* Allocates the activation buffers (zeroed), builds the TMA descriptors and the layer program. */
int c4a0_net_create(const c4a0_net_spec* spec, c4a0_net** out);
void c4a0_net_destroy(c4a0_net* net);
size_t c4a0_net_device_bytes(const c4a0_net* net);

/* Device address of activation buffer `buffer` ([max_rows rounded up to 128][buffer_cols] bf16). */
int c4a0_net_buffer(c4a0_net* net, uint32_t buffer, void** base_dev, uint32_t* cols, uint32_t* rows);

/* Where the answers go: logits [max_rows][7], q_penalty [max_rows], q_no_penalty [max_rows], f32
 * (the engine's c4a0_engine_bind_io buffers). */
int c4a0_net_bind_outputs(c4a0_net* net, float* logits_dev, float* q_penalty_dev, float* q_no_penalty_dev);

/* Rows of a forward pass read on the device when the kernel starts: max(*a, *b) (both u32; pass the
 * engine's pair from c4a0_engine_rows_count_dev).  Overrides the `rows` argument of forward(). */
int c4a0_net_bind_row_count(c4a0_net* net, const uint32_t* a_dev, const uint32_t* b_dev);

/* Enqueue one forward pass over rows [0, rows) (or the bound device-side count).  One kernel launch;
 * safe to capture into a CUDA graph. */
int c4a0_net_forward(c4a0_net* net, uint32_t rows, void* stream);
/* forward() with launch flags.  C4A0_NET_LAUNCH_PDL: programmatic dependent launch — the kernel's CTAs may become
 * resident and set up shared / tensor memory while the preceding kernel of `stream` is still draining; everything
 * that reads memory waits (griddepcontrol.wait) until that kernel has completed.  Used by c4a0_engine_run_net. */
#define C4A0_NET_LAUNCH_PDL 1u
int c4a0_net_forward_ex(c4a0_net* net, uint32_t rows, void* stream, uint32_t flags);

/* Diagnostics: one forward pass in which CTA `cta` logs (SM cycle counter << 8 | tag) events of its three
 * roles into out[3][4096] (producer, MMA issuer, first epilogue thread; tags: 1 role start, 2 tile start,
 * 3 dependency met, 4 barrier wait over, 5 K step issued, 6 accumulator ready, 7 accumulator released,
 * 8 tile published; 0 = unused entry).  Synchronises the stream. */
int c4a0_net_debug_trace(c4a0_net* net, uint32_t rows, uint32_t cta, void* stream, uint64_t* out, size_t n_out);

/* forward() bracketed by CUDA events; synchronises the stream. */
int c4a0_net_forward_timed(c4a0_net* net, uint32_t rows, void* stream, float* ms);

#ifdef __cplusplus
}
#endif
#endif
