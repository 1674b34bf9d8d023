// c4_rules.cuh — Connect-Four bitboard rules, shared by host C++ and sm_100a device code.
//
// Semantics follow the reference's rust/src/c4r.rs bit for bit (layout: bit = row*7 + col,
// row 0 = bottom, 42 bits used; `value` marks the stones of the side to move), but nothing here
// is a translation of its loops: every operation is a branch-free mask expression.
//   make_move / invert   c4r.rs:58-72, 125-129
//   legal_moves          c4r.rs:266-269
//   is_terminal_state    c4r.rs:165-249   (69 win masks == four shift-and tests, proved equal in
//                                          tests/test_oracle_rules.py, tests/test_math_host.py (host) and tests/test_gpu_engine.py (device) against the oracle's 69 masks)
//   terminal value       c4r.rs:253-263
//   NN planes            c4r.rs:378-392
//   flip_h               c4r.rs:289-299
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define C4_HD __host__ __device__ __forceinline__
#else
#define C4_HD inline
#endif

namespace c4 {

constexpr int N_ROWS = 6, N_COLS = 7, N_CELLS = 42, PLANE_LEN = 84;
constexpr uint64_t BOARD = (1ull << 42) - 1;
constexpr uint64_t COL0 = 0x0810204081ull;  // bits 0,7,14,21,28,35: column 0
// columns 0..3 / 3..6 of every row (anchors of horizontal and diagonal fours)
constexpr uint64_t COLS_0_3 = 0x0Full * COL0;
constexpr uint64_t COLS_3_6 = 0x78ull * COL0;

enum Terminal : int { NONE = 0, PLAYER_WIN = 1, OPPONENT_WIN = 2, DRAW = 3 };

struct Pos {
  uint64_t mask, value;
};

C4_HD int popc64(uint64_t x) {
#if defined(__CUDA_ARCH__)
  return __popcll(x);
#else
  return __builtin_popcountll(x);
#endif
}

C4_HD int ply(uint64_t mask) { return popc64(mask); }

// bit c set <=> column c still has room (top-row cell empty)
C4_HD unsigned legal_mask(uint64_t mask) { return (unsigned)((~mask >> 35) & 0x7full); }

// Drop a stone of the side to move into `col` (must be legal), then swap sides.
C4_HD Pos make_move(Pos p, int col) {
  uint64_t empty = (COL0 << col) & ~p.mask;
  uint64_t bit = empty & (0 - empty);  // lowest empty cell of the column
  Pos r;
  r.mask = p.mask | bit;
  r.value = ~(p.value | bit) & r.mask;
  return r;
}

C4_HD bool has_four(uint64_t t) {
  uint64_t v = t & (t >> 7) & (t >> 14) & (t >> 21);                 // vertical
  uint64_t h = t & (t >> 1) & (t >> 2) & (t >> 3) & COLS_0_3;        // horizontal
  uint64_t d = t & (t >> 8) & (t >> 16) & (t >> 24) & COLS_0_3;      // diagonal up-right
  uint64_t a = t & (t >> 6) & (t >> 12) & (t >> 18) & COLS_3_6;      // diagonal up-left
  return (v | h | d | a) != 0;
}

// Order of checks as in c4r.rs:228-238: side to move, then opponent, then full board.
C4_HD int terminal_state(Pos p) {
  if (has_four(p.mask & p.value)) return PLAYER_WIN;
  if (has_four(p.mask & ~p.value)) return OPPONENT_WIN;
  if (p.mask == BOARD) return DRAW;
  return NONE;
}

// c4r.rs:253-263.  Plain f32 ops in the reference's order (compile with -fmad=false).
C4_HD int terminal_value(Pos p, float c_ply_penalty, float* qp, float* qn) {
  int t = terminal_state(p);
  float m = c_ply_penalty * (float)ply(p.mask);
  if (t == PLAYER_WIN) {
    *qp = 1.0f - m;
    *qn = 1.0f;
  } else if (t == OPPONENT_WIN) {
    *qp = -1.0f + m;
    *qn = -1.0f;
  } else {
    *qp = 0.0f;
    *qn = 0.0f;
  }
  return t;
}

// element i (0..83) of the [2][6][7] input planes: channel 0 = side to move, 1 = opponent
C4_HD float plane_elem(Pos p, int i) {
  uint64_t bits = (i < N_CELLS) ? (p.mask & p.value) : (p.mask & ~p.value);
  int b = (i < N_CELLS) ? i : i - N_CELLS;
  return (float)((bits >> b) & 1ull);
}

C4_HD uint64_t flip_rows(uint64_t x) {
  uint64_t r = 0;
#pragma unroll
  for (int c = 0; c < N_COLS; c++) r |= ((x >> c) & COL0) << (N_COLS - 1 - c);
  return r;
}
// 49 bits that identify a position (key of the evaluation cache, engine.cu): the side-to-move's
// stones plus one marker bit on the first empty cell of every column (row 6 for a full column).
// Stones obey gravity, so the marker is the highest set bit of its column and the key decodes
// uniquely: below the marker, 1 = side to move, 0 = opponent.
C4_HD uint64_t pos_key(Pos p) {
  const uint64_t rows7 = (1ull << 49) - 1ull;
  return p.value | ((((p.mask << 7) | 0x7full) & ~p.mask) & rows7);
}

C4_HD Pos flip_h(Pos p) {
  Pos r;
  r.mask = flip_rows(p.mask);
  r.value = flip_rows(p.value);
  return r;
}

}  // namespace c4
