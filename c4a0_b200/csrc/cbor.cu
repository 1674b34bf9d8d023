// cbor.cu — `PlayGamesResult` <-> the bytes serde_cbor 0.11.2 writes for it (host code only).
//
// The reference pickles a PlayGamesResult as CBOR (rust/src/pybridge.rs:73-92: `serde_cbor::to_vec`), i.e.
//   {"results": [ {"metadata": {"game_id", "player0_id", "player1_id"},
//                  "samples": [ {"pos": {"mask", "value"}, "policy": [f32; 7], "q_penalty", "q_no_penalty"} ]} ]}
// with serde's conventions: structs as definite-length maps keyed by field name in declaration order,
// unsigned integers with the shortest head, and every f32 written as a half float when that is lossless,
// else as a single (serde_cbor's `serialize_f32`).  A job of 131,072 games is 2.6 M samples; this is the
// bulk path (c4a0_rust/_cbor.py is the same codec in Python, used by the tests as the cross-check).
#include <stdint.h>
#include <string.h>

#include <string>
#include <vector>

#include "common.cuh"

namespace {

using c4host::fail;
constexpr int MAXS = C4A0_MAX_SAMPLES;

struct Writer {
  uint8_t* p;
  size_t cap, n = 0;
  void put(uint8_t b) {
    if (n < cap) p[n] = b;
    n++;
  }
  void raw(const void* src, size_t len) {
    if (n + len <= cap) memcpy(p + n, src, len);
    n += len;
  }
  void head(int major, uint64_t v) {
    const uint8_t m = (uint8_t)(major << 5);
    if (v < 24) {
      put(m | (uint8_t)v);
    } else if (v < (1ull << 8)) {
      put(m | 24);
      put((uint8_t)v);
    } else if (v < (1ull << 16)) {
      put(m | 25);
      put((uint8_t)(v >> 8));
      put((uint8_t)v);
    } else if (v < (1ull << 32)) {
      put(m | 26);
      for (int s = 24; s >= 0; s -= 8) put((uint8_t)(v >> s));
    } else {
      put(m | 27);
      for (int s = 56; s >= 0; s -= 8) put((uint8_t)(v >> s));
    }
  }
  void text(const char* s) {
    const size_t len = strlen(s);
    head(3, len);
    raw(s, len);
  }
  // serde_cbor serialize_f32: half when lossless (incl. +-0, +-inf), NaN as f9 7e00, else single
  void f32(float f) {
    uint32_t b;
    memcpy(&b, &f, 4);
    const uint32_t sign = b >> 31, ex = (b >> 23) & 0xffu, man = b & 0x7fffffu;
    uint16_t h = 0;
    bool half = false;
    if (ex == 0xffu) {
      if (man) {
        put(0xf9);
        put(0x7e);
        put(0x00);
        return;
      }
      h = (uint16_t)((sign << 15) | 0x7c00u);
      half = true;
    } else if (ex == 0 && man == 0) {
      h = (uint16_t)(sign << 15);
      half = true;
    } else if (ex != 0) {
      const int e = (int)ex - 127;
      if (e >= -14 && e <= 15) {
        if ((man & 0x1fffu) == 0) {
          h = (uint16_t)((sign << 15) | ((uint32_t)(e + 15) << 10) | (man >> 13));
          half = true;
        }
      } else if (e >= -24 && e < -14) {  // a half subnormal: m * 2^-24
        const uint32_t full = (1u << 23) | man;
        const int sh = -(e + 1);  // 14..23
        if ((full & ((1u << sh) - 1u)) == 0) {
          h = (uint16_t)((sign << 15) | (full >> sh));
          half = true;
        }
      }
    }
    if (half) {
      put(0xf9);
      put((uint8_t)(h >> 8));
      put((uint8_t)h);
    } else {
      put(0xfa);
      for (int s = 24; s >= 0; s -= 8) put((uint8_t)(b >> s));
    }
  }
};

struct Reader {
  const uint8_t* p;
  size_t len, i = 0;
  bool ok = true;
  uint8_t byte() {
    if (i >= len) {
      ok = false;
      return 0;
    }
    return p[i++];
  }
  // major type + argument (definite lengths only)
  bool head(int* major, uint64_t* v, uint8_t* info_out = nullptr) {
    const uint8_t ib = byte();
    if (!ok) return false;
    *major = ib >> 5;
    const uint8_t info = ib & 31;
    if (info_out) *info_out = info;
    if (info < 24) {
      *v = info;
    } else if (info <= 27) {
      const int nb = 1 << (info - 24);
      uint64_t x = 0;
      for (int k = 0; k < nb; k++) x = (x << 8) | byte();
      *v = x;
    } else {
      ok = false;
    }
    return ok;
  }
  bool expect(int major, uint64_t* v) {
    int m;
    return head(&m, v) && (m == major || (ok = false));
  }
  bool key(std::string* s) {
    uint64_t n;
    if (!expect(3, &n) || i + n > len) return ok = false;
    s->assign(reinterpret_cast<const char*>(p + i), (size_t)n);
    i += (size_t)n;
    return true;
  }
  bool uint(uint64_t* v) { return expect(0, v); }
  bool f32(float* out) {
    int m;
    uint64_t v;
    uint8_t info;
    if (!head(&m, &v, &info)) return false;
    if (m == 7 && info == 25) {  // half
      const uint32_t h = (uint32_t)v, sign = h >> 15, ex = (h >> 10) & 31u, man = h & 0x3ffu;
      uint32_t b;
      if (ex == 31) {
        b = (sign << 31) | 0x7f800000u | (man << 13);
      } else if (ex == 0) {
        if (man == 0) {
          b = sign << 31;
        } else {  // subnormal half -> normal single
          int e = -14;
          uint32_t mm = man;
          while (!(mm & 0x400u)) {
            mm <<= 1;
            e--;
          }
          b = (sign << 31) | ((uint32_t)(e + 127) << 23) | ((mm & 0x3ffu) << 13);
        }
      } else {
        b = (sign << 31) | ((ex - 15 + 127) << 23) | (man << 13);
      }
      memcpy(out, &b, 4);
      return true;
    }
    if (m == 7 && info == 26) {
      const uint32_t b = (uint32_t)v;
      memcpy(out, &b, 4);
      return true;
    }
    if (m == 7 && info == 27) {
      double d;
      memcpy(&d, &v, 8);
      *out = (float)d;
      return true;
    }
    if (m == 0) {  // an integer where a float is expected (other encoders)
      *out = (float)v;
      return true;
    }
    return ok = false;
  }
  // skip one item of any type
  bool skip() {
    int m;
    uint64_t v;
    if (!head(&m, &v)) return false;
    if (m == 2 || m == 3) {
      if (i + v > len) return ok = false;
      i += (size_t)v;
    } else if (m == 4) {
      for (uint64_t k = 0; k < v && ok; k++) skip();
    } else if (m == 5) {
      for (uint64_t k = 0; k < 2 * v && ok; k++) skip();
    } else if (m == 6) {
      skip();
    }
    return ok;
  }
};

}  // namespace

extern "C" {

int c4a0_results_to_cbor(const uint64_t* meta, const uint32_t* n_samples, const uint64_t* mask, const uint64_t* value,
                         const float* policy, const float* qp, const float* qn, uint32_t n_games, uint8_t* out, size_t cap,
                         size_t* needed) {
  if (!needed || (n_games && (!meta || !n_samples || !mask || !value || !policy || !qp || !qn)))
    return fail(C4A0_E_INVALID, "null argument");
  Writer w{out, out ? cap : 0};
  w.head(5, 1);
  w.text("results");
  w.head(4, n_games);
  for (uint32_t g = 0; g < n_games; g++) {
    if (n_samples[g] > (uint32_t)MAXS) return fail(C4A0_E_INVALID, "game %u has more than %d samples", g, MAXS);
    w.head(5, 2);
    w.text("metadata");
    w.head(5, 3);
    w.text("game_id");
    w.head(0, meta[3 * (size_t)g]);
    w.text("player0_id");
    w.head(0, meta[3 * (size_t)g + 1]);
    w.text("player1_id");
    w.head(0, meta[3 * (size_t)g + 2]);
    w.text("samples");
    w.head(4, n_samples[g]);
    for (uint32_t k = 0; k < n_samples[g]; k++) {
      const size_t s = (size_t)g * MAXS + k;
      w.head(5, 4);
      w.text("pos");
      w.head(5, 2);
      w.text("mask");
      w.head(0, mask[s]);
      w.text("value");
      w.head(0, value[s]);
      w.text("policy");
      w.head(4, 7);
      for (int c = 0; c < 7; c++) w.f32(policy[s * 7 + c]);
      w.text("q_penalty");
      w.f32(qp[s]);
      w.text("q_no_penalty");
      w.f32(qn[s]);
    }
  }
  *needed = w.n;
  return 0;
}

// Two calls: with null arrays only *n_games is returned; then with arrays of n_games ([G][3], [G], [G][43], ...).
int c4a0_results_from_cbor(const uint8_t* buf, size_t len, uint32_t* n_games, uint64_t* meta, uint32_t* n_samples,
                           uint64_t* mask, uint64_t* value, float* policy, float* qp, float* qn) {
  if (!buf || !n_games) return fail(C4A0_E_INVALID, "null argument");
  const bool fill = meta != nullptr;
  if (fill && (!n_samples || !mask || !value || !policy || !qp || !qn)) return fail(C4A0_E_INVALID, "null argument");
  Reader r{buf, len};
  std::string k;
  uint64_t n_top = 0, games = 0;
  bool have_results = false;
  if (!r.expect(5, &n_top)) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: not a map");
  for (uint64_t t = 0; t < n_top && r.ok; t++) {
    if (!r.key(&k)) break;
    if (k != "results") {
      r.skip();
      continue;
    }
    have_results = true;
    if (!r.expect(4, &games)) break;
    if (fill && games != *n_games) return fail(C4A0_E_INVALID, "game count changed between the two calls");
    for (uint64_t g = 0; g < games && r.ok; g++) {
      uint64_t nf = 0, ns = 0;
      bool have_md = false, have_samples = false;
      if (!r.expect(5, &nf)) break;
      for (uint64_t f = 0; f < nf && r.ok; f++) {
        if (!r.key(&k)) break;
        if (k == "metadata") {
          uint64_t nm = 0, seen = 0;
          if (!r.expect(5, &nm)) break;
          for (uint64_t q = 0; q < nm && r.ok; q++) {
            uint64_t v = 0;
            if (!r.key(&k)) break;
            const int idx = k == "game_id" ? 0 : k == "player0_id" ? 1 : k == "player1_id" ? 2 : -1;
            if (idx < 0) {
              r.skip();
              continue;
            }
            if (!r.uint(&v)) break;
            seen |= 1u << idx;
            if (fill) meta[3 * g + idx] = v;
          }
          have_md = seen == 7;
        } else if (k == "samples") {
          if (!r.expect(4, &ns)) break;
          if (ns > (uint64_t)MAXS) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: a game cannot have more than %d samples", MAXS);
          have_samples = true;
          if (fill) n_samples[g] = (uint32_t)ns;
          for (uint64_t s = 0; s < ns && r.ok; s++) {
            const size_t o = (size_t)g * MAXS + s;
            uint64_t nq = 0;
            unsigned seen = 0;
            if (!r.expect(5, &nq)) break;
            for (uint64_t q = 0; q < nq && r.ok; q++) {
              if (!r.key(&k)) break;
              if (k == "pos") {
                uint64_t np = 0;
                if (!r.expect(5, &np)) break;
                for (uint64_t z = 0; z < np && r.ok; z++) {
                  uint64_t v = 0;
                  if (!r.key(&k)) break;
                  if (k == "mask") {
                    if (r.uint(&v)) seen |= 1;
                    if (fill) mask[o] = v;
                  } else if (k == "value") {
                    if (r.uint(&v)) seen |= 2;
                    if (fill) value[o] = v;
                  } else {
                    r.skip();
                  }
                }
              } else if (k == "policy") {
                uint64_t np = 0;
                if (!r.expect(4, &np) || np != 7) {
                  r.ok = false;
                  break;
                }
                for (int c = 0; c < 7 && r.ok; c++) {
                  float x = 0;
                  r.f32(&x);
                  if (fill) policy[o * 7 + c] = x;
                }
                seen |= 4;
              } else if (k == "q_penalty") {
                float x = 0;
                if (r.f32(&x)) seen |= 8;
                if (fill) qp[o] = x;
              } else if (k == "q_no_penalty") {
                float x = 0;
                if (r.f32(&x)) seen |= 16;
                if (fill) qn[o] = x;
              } else {
                r.skip();
              }
            }
            if (r.ok && seen != 31) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: a sample lacks a field");
          }
        } else {
          r.skip();
        }
      }
      if (r.ok && !(have_md && have_samples)) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: a game lacks metadata or samples");
    }
  }
  if (!r.ok) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: malformed or truncated at byte %zu", r.i);
  if (!have_results) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: no 'results' key");
  if (r.i != len) return fail(C4A0_E_INVALID, "invalid PlayGamesResult CBOR: trailing bytes");
  *n_games = (uint32_t)games;
  return 0;
}

}  // extern "C"
