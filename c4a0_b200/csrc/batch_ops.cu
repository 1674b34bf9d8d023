// batch_ops.cu — stand-alone batch kernels (rules / math / sampling) and the host builds of the
// shared math.  These entry points exist so that the parity tests can exercise every device
// function of c4_rules.cuh / c4_math.cuh / c4_rng.cuh in isolation, on arbitrary inputs, through
// the same C-ABI the engine uses.
#include <stdarg.h>
#include <stdio.h>

#include <string>

#include "c4_math.cuh"
#include "c4_rng.cuh"
#include "c4_rules.cuh"
#include <cuda_bf16.h>

#include "common.cuh"

namespace c4host {
static thread_local std::string g_err;
int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}
const char* last_error() { return g_err.c_str(); }
int no_gpu_error() {
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(C4A0_E_CUDA, "no CUDA device available (%s): the engine has no CPU fallback",
                err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
  }
  return 0;
}
}  // namespace c4host

namespace {
using c4::Pos;
using c4host::blocks_for;
using c4host::DevBuf;
using c4host::fail;
using c4host::no_gpu_error;

// ------------------------------------------------------------------------------------------------
// Stand-alone batch kernels
// ------------------------------------------------------------------------------------------------
__global__ void k_rules(const uint64_t* mask, const uint64_t* value, size_t n, float c_pen,
                        int32_t* terminal, uint32_t* legal, int32_t* ply, float* qp, float* qn,
                        uint64_t* cm, uint64_t* cv, float* planes, uint64_t* fm, uint64_t* fv) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Pos p{mask[i], value[i]};
  float a, b;
  int t = c4::terminal_value(p, c_pen, &a, &b);
  if (terminal) terminal[i] = t;
  unsigned lg = c4::legal_mask(p.mask);
  if (legal) legal[i] = lg;
  if (ply) ply[i] = c4::ply(p.mask);
  if (qp) qp[i] = a;
  if (qn) qn[i] = b;
  if (cm && cv)
    for (int c = 0; c < 7; c++) {
      Pos ch{0ull, 0ull};
      if ((lg >> c) & 1u) ch = c4::make_move(p, c);
      cm[i * 7 + c] = ch.mask;
      cv[i * 7 + c] = ch.value;
    }
  if (planes)
    for (int k = 0; k < 84; k++) planes[i * 84 + k] = c4::plane_elem(p, k);
  if (fm && fv) {
    Pos f = c4::flip_h(p);
    fm[i] = f.mask;
    fv[i] = f.value;
  }
}
__global__ void k_math(int op, const float* in, float* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  out[i] = op == C4A0_MATH_LOGF ? c4::c4_logf(in[i]) : c4::c4_expf(in[i]);
}
__global__ void k_softmax(const float* logits, const uint32_t* legal, float* out, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float x[7], o[7];
  for (int k = 0; k < 7; k++) x[k] = ((legal[i] >> k) & 1u) ? logits[i * 7 + k] : -c4::f32_inf();
  if (!c4::softmax7(x, o))
    for (int k = 0; k < 7; k++) o[k] = c4::f32_nan();
  for (int k = 0; k < 7; k++) out[i * 7 + k] = o[k];
}
__global__ void k_sample(const float* policy, const float* temperature, const uint64_t* seed,
                         float* tempered, int32_t* column, size_t n) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float p[7], t[7];
  for (int k = 0; k < 7; k++) p[k] = policy[i * 7 + k];
  c4::apply_temperature7(p, temperature[i], t);
  for (int k = 0; k < 7; k++) tempered[i * 7 + k] = t[k];
  column[i] = c4::weighted_sample7(t, seed[i]);
}

// nn.py:116-130 — log_softmax(policy), tanh(values), written where the engine reads them
template <typename T>
__device__ __forceinline__ float head_ld(const T* p);
template <>
__device__ __forceinline__ float head_ld<float>(const float* p) { return *p; }
template <>
__device__ __forceinline__ float head_ld<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }
template <typename T>
__global__ void k_head_epilogue(const T* __restrict__ pol, const T* __restrict__ val, uint32_t ldp, uint32_t ldv,
                                uint32_t rows, float* __restrict__ logits, float* __restrict__ qp, float* __restrict__ qn) {
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float x[7];
  float mx = -c4::f32_inf();
#pragma unroll
  for (int k = 0; k < 7; k++) {
    x[k] = head_ld(pol + (size_t)r * ldp + k);
    mx = fmaxf(mx, x[k]);
  }
  float s = 0.0f;
#pragma unroll
  for (int k = 0; k < 7; k++) s += expf(x[k] - mx);
  const float lse = mx + logf(s);
#pragma unroll
  for (int k = 0; k < 7; k++) logits[(size_t)r * 7 + k] = x[k] - lse;
  qp[r] = tanhf(head_ld(val + (size_t)r * ldv));
  qn[r] = tanhf(head_ld(val + (size_t)r * ldv + 1));
}

// nn.py:100-130 — both output layers and the output stage, one warp per row.  The 9 weight rows live in
// shared memory as f32; a lane owns the 8-element chunks lane, lane + 32, ... of the row.
template <typename T>
__device__ __forceinline__ void load8(const T* p, float (&v)[8]);
template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&v)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
  v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&v)[8]) {
  const uint4 u = *reinterpret_cast<const uint4*>(p);
  const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    v[2 * i] = __uint_as_float(w[i] << 16);
    v[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
  }
}
constexpr int HEADS_WARPS = 8;
template <typename T, int HEADS_ROWS>
__global__ void __launch_bounds__(HEADS_WARPS * 32) k_heads(const T* __restrict__ hp, const T* __restrict__ hv, uint32_t ldp,
                                                            uint32_t ldv, uint32_t F, const T* __restrict__ wp_t,
                                                            const float* __restrict__ bp, const T* __restrict__ wv_t,
                                                            const float* __restrict__ bv, uint32_t rows,
                                                            float* __restrict__ logits, float* __restrict__ qp,
                                                            float* __restrict__ qn) {
  // [9][2][F/2]: 7 policy rows, then 2 value rows; elements 8c..8c+3 of a row sit at [0][4c..], elements
  // 8c+4..8c+7 at [1][4c..], so that consecutive lanes read consecutive 16-byte vectors (no bank conflicts)
  extern __shared__ float sh_w[];
  const uint32_t H = F >> 1;
  const uint32_t chunks = F >> 3;
  for (uint32_t i = threadIdx.x; i < 9u * chunks; i += blockDim.x) {  // one 8-element chunk per step
    const uint32_t o = i / chunks, c = i - o * chunks;
    float w[8];
    load8<T>((o < 7 ? wp_t + (size_t)o * F : wv_t + (size_t)(o - 7) * F) + 8u * c, w);
    *reinterpret_cast<float4*>(sh_w + (size_t)o * F + 4u * c) = make_float4(w[0], w[1], w[2], w[3]);
    *reinterpret_cast<float4*>(sh_w + (size_t)o * F + H + 4u * c) = make_float4(w[4], w[5], w[6], w[7]);
  }
  __syncthreads();
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  // A warp takes HEADS_ROWS rows at a time, so that a weight vector read from shared memory is used for
  // all of them (shared-memory bandwidth, not HBM, bounds a one-row-per-warp version) and eight global
  // loads per lane are in flight.
  for (uint32_t r0 = (blockIdx.x * HEADS_WARPS + warp) * HEADS_ROWS; r0 < rows; r0 += gridDim.x * HEADS_WARPS * HEADS_ROWS) {
    float acc[HEADS_ROWS][9];
#pragma unroll
    for (int q = 0; q < HEADS_ROWS; q++)
#pragma unroll
      for (int o = 0; o < 9; o++) acc[q][o] = 0.0f;
    for (uint32_t c = lane; c < chunks; c += 32u) {
      float a[HEADS_ROWS][8], b[HEADS_ROWS][8];
#pragma unroll
      for (int q = 0; q < HEADS_ROWS; q++) {
        const uint32_t r = r0 + q < rows ? r0 + q : rows - 1u;  // a clamped duplicate instead of a branch
        load8<T>(hp + (size_t)r * ldp + 8u * c, a[q]);
        load8<T>(hv + (size_t)r * ldv + 8u * c, b[q]);
      }
#pragma unroll
      for (int o = 0; o < 9; o++) {
        const float4 w0 = *reinterpret_cast<const float4*>(sh_w + (size_t)o * F + 4u * c);
        const float4 w1 = *reinterpret_cast<const float4*>(sh_w + (size_t)o * F + H + 4u * c);
#pragma unroll
        for (int q = 0; q < HEADS_ROWS; q++) {
          const float(&v)[8] = o < 7 ? a[q] : b[q];
          acc[q][o] = fmaf(v[0], w0.x, fmaf(v[1], w0.y, fmaf(v[2], w0.z, fmaf(v[3], w0.w, fmaf(v[4], w1.x, fmaf(v[5], w1.y, fmaf(v[6], w1.z, fmaf(v[7], w1.w, acc[q][o]))))))));
        }
      }
    }
#pragma unroll
    for (int q = 0; q < HEADS_ROWS; q++) {
#pragma unroll
      for (int o = 0; o < 9; o++) {
#pragma unroll
        for (int m = 16; m >= 1; m >>= 1) acc[q][o] += __shfl_xor_sync(0xffffffffu, acc[q][o], m);
      }
      // output stage across the lanes: lane k < 7 owns policy column k, lanes 7 and 8 the two values
      float v = acc[q][0];
#pragma unroll
      for (int k = 1; k < 9; k++) v = lane == (uint32_t)k ? acc[q][k] : v;
      const float x = lane < 7u ? v + bp[lane] : -c4::f32_inf();
      float mx = x;
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, m));  // lanes 0..7
      mx = __shfl_sync(0xffffffffu, mx, 0);
      float e = lane < 7u ? expf(x - mx) : 0.0f;
#pragma unroll
      for (int m = 4; m >= 1; m >>= 1) e += __shfl_xor_sync(0xffffffffu, e, m);
      const float lse = mx + logf(__shfl_sync(0xffffffffu, e, 0));
      const uint32_t r = r0 + q;
      if (r < rows) {
        if (lane < 7u) logits[(size_t)r * 7 + lane] = x - lse;
        if (lane == 7u) qp[r] = tanhf(v + bv[0]);
        if (lane == 8u) qn[r] = tanhf(v + bv[1]);
      }
    }
  }
}

}  // namespace

extern "C" {

int c4a0_heads(const void* hp, const void* hv, uint32_t dtype, uint32_t ld_hp, uint32_t ld_hv, uint32_t F, const void* wp_t,
               const float* bp, const void* wv_t, const float* bv, uint32_t rows, float* logits, float* qp, float* qn,
               void* stream) {
  if (!hp || !hv || !wp_t || !bp || !wv_t || !bv || !logits || !qp || !qn) return fail(C4A0_E_INVALID, "null argument");
  if (dtype > C4A0_PLANES_BF16 || F == 0 || (F & 7u) || ld_hp < F || ld_hv < F || (ld_hp & 7u) || (ld_hv & 7u))
    return fail(C4A0_E_INVALID, "bad head layout: F and the row strides must be multiples of 8");
  if ((((uintptr_t)hp | (uintptr_t)hv) & 15u) != 0) return fail(C4A0_E_INVALID, "activations must be 16-byte aligned");
  if (rows == 0) return 0;
  const size_t smem = (size_t)9 * F * sizeof(float);
  if (smem > 200 * 1024) return fail(C4A0_E_INVALID, "F too large for c4a0_heads");
  cudaStream_t s = (cudaStream_t)stream;
  // small batches: one row per warp, spread over the SMs; large ones: four rows per warp (weight reuse)
  const bool wide = rows > 148u * 2u * HEADS_WARPS;
  const unsigned per_cta = HEADS_WARPS * (wide ? 4 : 1);
  unsigned grid = (rows + per_cta - 1) / per_cta;
  if (grid > 148u * 4u) grid = 148u * 4u;  // persistent: each CTA stages the weights once and strides over the rows
#define C4A0_LAUNCH_HEADS(T, R)                                                                                        \
  do {                                                                                                                 \
    auto k = k_heads<T, R>;                                                                                            \
    if (smem > 48 * 1024) CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));         \
    k<<<grid, HEADS_WARPS * 32, smem, s>>>((const T*)hp, (const T*)hv, ld_hp, ld_hv, F, (const T*)wp_t, bp, (const T*)wv_t, \
                                           bv, rows, logits, qp, qn);                                                  \
  } while (0)
  if (dtype == C4A0_PLANES_BF16) {
    if (wide) C4A0_LAUNCH_HEADS(__nv_bfloat16, 4); else C4A0_LAUNCH_HEADS(__nv_bfloat16, 1);
  } else {
    if (wide) C4A0_LAUNCH_HEADS(float, 4); else C4A0_LAUNCH_HEADS(float, 1);
  }
#undef C4A0_LAUNCH_HEADS
  CK(cudaGetLastError());
  return 0;
}

int c4a0_head_epilogue(const void* policy_head, const void* value_head, uint32_t dtype, uint32_t ld_policy,
                       uint32_t ld_value, uint32_t rows, float* logits, float* qp, float* qn, void* stream) {
  if (!policy_head || !value_head || !logits || !qp || !qn) return fail(C4A0_E_INVALID, "null argument");
  if (dtype > C4A0_PLANES_BF16 || ld_policy < 7 || ld_value < 2) return fail(C4A0_E_INVALID, "bad head layout");
  if (rows == 0) return 0;
  cudaStream_t s = (cudaStream_t)stream;
  const unsigned grid = (rows + 127u) / 128u;
  if (dtype == C4A0_PLANES_BF16)
    k_head_epilogue<__nv_bfloat16><<<grid, 128, 0, s>>>((const __nv_bfloat16*)policy_head, (const __nv_bfloat16*)value_head,
                                                        ld_policy, ld_value, rows, logits, qp, qn);
  else
    k_head_epilogue<float><<<grid, 128, 0, s>>>((const float*)policy_head, (const float*)value_head, ld_policy, ld_value,
                                                rows, logits, qp, qn);
  CK(cudaGetLastError());
  return 0;
}

const char* c4a0_last_error(void) { return c4host::last_error(); }
int c4a0_abi_version(void) { return C4A0_ABI_VERSION; }

// ---- stand-alone batch entry points --------------------------------------------------------------
#define BATCH_PROLOGUE()            \
  do {                              \
    int _r = no_gpu_error();        \
    if (_r) return _r;              \
    CK(cudaSetDevice(device));      \
  } while (0)
#define UP(dbuf, hptr, count)                                                              \
  do {                                                                                     \
    CK((dbuf).alloc(count));                                                               \
    CK(cudaMemcpy((dbuf).p, hptr, (count) * sizeof(*(dbuf).p), cudaMemcpyHostToDevice));   \
  } while (0)
#define DOWN(hptr, dbuf, count) \
  CK(cudaMemcpy(hptr, (dbuf).p, (count) * sizeof(*(dbuf).p), cudaMemcpyDeviceToHost))

int c4a0_rules_batch(int device, const uint64_t* mask, const uint64_t* value, size_t n, float c_pen,
                     int32_t* terminal, uint32_t* legal, int32_t* ply, float* qp, float* qn,
                     uint64_t* cm, uint64_t* cv, float* planes, uint64_t* fm, uint64_t* fv) {
  if (!mask || !value) return fail(C4A0_E_INVALID, "null positions");
  BATCH_PROLOGUE();
  if (n == 0) return 0;
  DevBuf<uint64_t> dm, dv, dcm, dcv, dfm, dfv;
  DevBuf<int32_t> dt, dp;
  DevBuf<uint32_t> dl;
  DevBuf<float> dqp, dqn, dpl;
  UP(dm, mask, n);
  UP(dv, value, n);
  if (terminal) CK(dt.alloc(n));
  if (legal) CK(dl.alloc(n));
  if (ply) CK(dp.alloc(n));
  if (qp) CK(dqp.alloc(n));
  if (qn) CK(dqn.alloc(n));
  if (cm && cv) { CK(dcm.alloc(n * 7)); CK(dcv.alloc(n * 7)); }
  if (planes) CK(dpl.alloc(n * 84));
  if (fm && fv) { CK(dfm.alloc(n)); CK(dfv.alloc(n)); }
  k_rules<<<blocks_for(n, 256), 256>>>(dm.p, dv.p, n, c_pen, dt.p, dl.p, dp.p, dqp.p, dqn.p, dcm.p, dcv.p,
                                       dpl.p, dfm.p, dfv.p);
  CK(cudaGetLastError());
  if (terminal) DOWN(terminal, dt, n);
  if (legal) DOWN(legal, dl, n);
  if (ply) DOWN(ply, dp, n);
  if (qp) DOWN(qp, dqp, n);
  if (qn) DOWN(qn, dqn, n);
  if (cm && cv) { DOWN(cm, dcm, n * 7); DOWN(cv, dcv, n * 7); }
  if (planes) DOWN(planes, dpl, n * 84);
  if (fm && fv) { DOWN(fm, dfm, n); DOWN(fv, dfv, n); }
  return 0;
}

int c4a0_math_batch(int device, int op, const float* in, float* out, size_t n) {
  if (!in || !out) return fail(C4A0_E_INVALID, "null argument");
  if (op != C4A0_MATH_LOGF && op != C4A0_MATH_EXPF) return fail(C4A0_E_INVALID, "bad op");
  BATCH_PROLOGUE();
  if (n == 0) return 0;
  DevBuf<float> di, dout;
  UP(di, in, n);
  CK(dout.alloc(n));
  k_math<<<blocks_for(n, 256), 256>>>(op, di.p, dout.p, n);
  CK(cudaGetLastError());
  DOWN(out, dout, n);
  return 0;
}

int c4a0_softmax_batch(int device, const float* logits, const uint32_t* legal, float* out, size_t n) {
  if (!logits || !legal || !out) return fail(C4A0_E_INVALID, "null argument");
  BATCH_PROLOGUE();
  if (n == 0) return 0;
  DevBuf<float> dl, dout;
  DevBuf<uint32_t> dg;
  UP(dl, logits, n * 7);
  UP(dg, legal, n);
  CK(dout.alloc(n * 7));
  k_softmax<<<blocks_for(n, 128), 128>>>(dl.p, dg.p, dout.p, n);
  CK(cudaGetLastError());
  DOWN(out, dout, n * 7);
  return 0;
}

int c4a0_sample_batch(int device, const float* policy, const float* temperature, const uint64_t* seed,
                      float* tempered, int32_t* column, size_t n) {
  if (!policy || !temperature || !seed || !tempered || !column) return fail(C4A0_E_INVALID, "null argument");
  BATCH_PROLOGUE();
  if (n == 0) return 0;
  DevBuf<float> dp, dt, dtemp;
  DevBuf<uint64_t> ds;
  DevBuf<int32_t> dc;
  UP(dp, policy, n * 7);
  UP(dt, temperature, n);
  UP(ds, seed, n);
  CK(dtemp.alloc(n * 7));
  CK(dc.alloc(n));
  k_sample<<<blocks_for(n, 128), 128>>>(dp.p, dt.p, ds.p, dtemp.p, dc.p, n);
  CK(cudaGetLastError());
  DOWN(tempered, dtemp, n * 7);
  DOWN(column, dc, n);
  return 0;
}

// ---- host builds of the shared math (CPU test-suite) ----------------------------------------------
void c4a0_host_logf(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = c4::c4_logf(in[i]);
}
void c4a0_host_expf(const float* in, float* out, size_t n) {
  for (size_t i = 0; i < n; i++) out[i] = c4::c4_expf(in[i]);
}
int c4a0_host_sample(const float* policy, float temperature, uint64_t seed, float* tempered) {
  float t[7];
  c4::apply_temperature7(policy, temperature, t);
  if (tempered)
    for (int i = 0; i < 7; i++) tempered[i] = t[i];
  return c4::weighted_sample7(t, seed);
}
int c4a0_host_terminal_state(uint64_t mask, uint64_t value) { return c4::terminal_state(Pos{mask, value}); }
void c4a0_host_make_move(uint64_t mask, uint64_t value, int col, uint64_t* om, uint64_t* ov) {
  Pos r = c4::make_move(Pos{mask, value}, col);
  *om = r.mask;
  *ov = r.value;
}

void c4a0_host_flip_h(uint64_t mask, uint64_t value, uint64_t* om, uint64_t* ov) {
  Pos r = c4::flip_h(Pos{mask, value});
  *om = r.mask;
  *ov = r.value;
}
uint64_t c4a0_host_pos_key(uint64_t mask, uint64_t value) { return c4::pos_key(Pos{mask, value}); }
void c4a0_host_shuffle(uint64_t seed, uint32_t* idx, size_t n) { c4::shuffle_indices(seed, idx, n); }
void c4a0_host_seed_to_key(uint64_t seed, uint32_t* key8) { c4::seed_to_key(seed, key8); }
void c4a0_host_stdrng_words(const uint32_t* key8, uint32_t* out, size_t n) {
  c4::StdRngStream rng(0);
  for (int i = 0; i < 8; i++) rng.key[i] = key8[i];
  for (size_t i = 0; i < n; i++) out[i] = rng.next_u32();
}

}  // extern "C"
