// c4_rng.cuh — seeded move sampling (rust/src/mcts.rs:214-222), host + device.
//
// The reference draws one column per move from a fresh `StdRng::seed_from_u64(game_id * (42 +
// n_moves))` through `WeightedIndex<f32>`.  Both come from the third-party crate rand 0.10.1
// (rust/Cargo.lock:1585-1591; StdRng = ChaCha12), which is not vendored under the reference and
// which no reference test pins — so this follows the crate's published algorithm:
//   seed_from_u64   : eight PCG32 (XSH-RR) outputs form the 256-bit ChaCha key
//   StdRng          : ChaCha, 12 rounds, 64-bit block counter 0, stream 0; first output word
//   WeightedIndex   : cumulative f32 left-fold; Uniform<f32>[0,total) from the top 23 bits;
//                     index = number of cumulative weights <= the draw
// Only the first 32-bit word of the first block is ever consumed per move.
#pragma once
#include <stdint.h>

#include "c4_math.cuh"

namespace c4 {

C4_HD uint32_t rotl32(uint32_t x, int n) { return (x << n) | (x >> (32 - n)); }

C4_HD void seed_to_key(uint64_t state, uint32_t key[8]) {
  const uint64_t MUL = 6364136223846793005ULL, INC = 11634580027462260723ULL;
  for (int i = 0; i < 8; i++) {
    state = state * MUL + INC;
    uint32_t xs = (uint32_t)(((state >> 18) ^ state) >> 27);
    uint32_t rot = (uint32_t)(state >> 59);
    key[i] = (xs >> rot) | (xs << ((32 - rot) & 31));
  }
}

#define C4_QR(a, b, c, d) \
  a += b; d ^= a; d = rotl32(d, 16); c += d; b ^= c; b = rotl32(b, 12); \
  a += b; d ^= a; d = rotl32(d, 8);  c += d; b ^= c; b = rotl32(b, 7);

// word 0 of ChaCha12 block 0 under `key` (counter = 0, stream = 0)
C4_HD uint32_t chacha12_first_word(const uint32_t key[8]) {
  uint32_t x0 = 0x61707865u, x1 = 0x3320646eu, x2 = 0x79622d32u, x3 = 0x6b206574u;
  uint32_t x4 = key[0], x5 = key[1], x6 = key[2], x7 = key[3];
  uint32_t x8 = key[4], x9 = key[5], x10 = key[6], x11 = key[7];
  uint32_t x12 = 0, x13 = 0, x14 = 0, x15 = 0;
  for (int i = 0; i < 6; i++) {
    C4_QR(x0, x4, x8, x12) C4_QR(x1, x5, x9, x13) C4_QR(x2, x6, x10, x14) C4_QR(x3, x7, x11, x15)
    C4_QR(x0, x5, x10, x15) C4_QR(x1, x6, x11, x12) C4_QR(x2, x7, x8, x13) C4_QR(x3, x4, x9, x14)
  }
  return x0 + 0x61707865u;
}

// WeightedIndex::new(w).sample(StdRng::seed_from_u64(seed)); -1 where the reference would panic
// (negative / non-finite weight, or zero total).
C4_HD int weighted_sample7(const float w[7], uint64_t seed) {
  float cumulative[6];
  float total = w[0];
  if (!(total >= 0.0f)) return -1;
  for (int i = 1; i < 7; i++) {
    if (!(w[i] >= 0.0f)) return -1;
    cumulative[i - 1] = total;
    total = total + w[i];
  }
  if (!(total > 0.0f) || total == f32_inf()) return -1;
  float scale = total;  // high - low with low = 0
  const float max_rand = bits_f32(0x3f7ffffeu);  // 1 - 2^-23: the largest value the 23-bit draw can take
  while (scale * max_rand + 0.0f >= total) scale = bits_f32(f32_bits(scale) - 1u);
  uint32_t key[8];
  seed_to_key(seed, key);
  uint32_t u = chacha12_first_word(key);
  float v01 = bits_f32((u >> 9) | 0x3f800000u) - 1.0f;
  float x = v01 * scale + 0.0f;
  int idx = 0;
  while (idx < 6 && cumulative[idx] <= x) idx++;
  return idx;
}

// ---- host only: StdRng stream + SliceRandom::shuffle (rust/src/pybridge.rs:110-113) -----------
// Same provenance note as above: rand 0.10.1's `shuffle` walks the slice upwards and swaps element
// i with a uniform index in [0, i]; the indices come from `IncreasingUniform`, which draws one u32
// below the largest product (i+1)(i+2)...(i+k) that fits 32 bits and peels k indices off it by
// division.  Bounded u32 draws are widening multiplies with one bias-correction draw.
struct StdRngStream {
  uint32_t key[8];
  uint32_t block[16];
  uint64_t counter = 0;
  int used = 16;
  explicit StdRngStream(uint64_t seed) { seed_to_key(seed, key); }
  void refill() {
    uint32_t in[16] = {0x61707865u, 0x3320646eu, 0x79622d32u, 0x6b206574u};
    for (int i = 0; i < 8; i++) in[4 + i] = key[i];
    in[12] = (uint32_t)counter;
    in[13] = (uint32_t)(counter >> 32);
    in[14] = in[15] = 0;
    uint32_t x[16];
    for (int i = 0; i < 16; i++) x[i] = in[i];
    for (int r = 0; r < 6; r++) {
      C4_QR(x[0], x[4], x[8], x[12]) C4_QR(x[1], x[5], x[9], x[13])
      C4_QR(x[2], x[6], x[10], x[14]) C4_QR(x[3], x[7], x[11], x[15])
      C4_QR(x[0], x[5], x[10], x[15]) C4_QR(x[1], x[6], x[11], x[12])
      C4_QR(x[2], x[7], x[8], x[13]) C4_QR(x[3], x[4], x[9], x[14])
    }
    for (int i = 0; i < 16; i++) block[i] = x[i] + in[i];
    counter++;
    used = 0;
  }
  uint32_t next_u32() {
    if (used >= 16) refill();
    return block[used++];
  }
  // uniform in [0, bound) — bound == 0 means the full 32-bit range
  uint32_t below(uint32_t bound) {
    if (bound == 0) return next_u32();
    uint64_t m = (uint64_t)next_u32() * bound;
    uint32_t hi = (uint32_t)(m >> 32), lo = (uint32_t)m;
    if (lo > (uint32_t)(0u - bound)) {
      uint32_t hi2 = (uint32_t)(((uint64_t)next_u32() * bound) >> 32);
      if ((uint32_t)(lo + hi2) < lo) hi++;
    }
    return hi;
  }
};

inline void shuffle_indices(uint64_t seed, uint32_t* idx, size_t n) {
  if (n < 2) return;
  StdRngStream rng(seed);
  uint32_t packed = 0;   // indices still packed in the last draw
  uint32_t left = 1;     // how many indices `packed` still holds (the first, for i = 0, is free)
  for (size_t i = 0; i < n; i++) {
    uint32_t span = (uint32_t)i + 1;  // choose in [0, i]
    if (left == 0) {
      // largest run span*(span+1)*...*(span+k-1) that still fits in 32 bits
      uint32_t prod = span, nxt = span + 1;
      while ((uint64_t)prod * nxt <= 0xffffffffULL) {
        prod *= nxt;
        nxt++;
      }
      packed = rng.below(prod);
      left = nxt - span;
    }
    left--;
    uint32_t j;
    if (left == 0) {
      j = packed;
    } else {
      j = packed % span;
      packed /= span;
    }
    uint32_t t = idx[i];
    idx[i] = idx[j];
    idx[j] = t;
  }
}

C4_HD uint64_t move_seed(uint64_t game_id, int n_moves) { return game_id * (uint64_t)(42 + n_moves); }

}  // namespace c4
