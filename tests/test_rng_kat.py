"""Known-answer tests that pin the restated `rand 0.10.1` machinery (rust/src/mcts.rs:214-222,
rust/src/pybridge.rs:110-113; rust/Cargo.lock:1585-1591) to PUBLIC vectors, in the oracle and in the product:

  * ChaCha with 12 / 20 / 8 rounds, all-zero 256-bit key and nonce, block 0 — the keystreams published in
    draft-strombergson-chacha-test-vectors (TC1) and RFC 7539 §2.3.2-style vectors;
  * `StdRng` value stability — the rand crate's own `test_stdrng_construction` (rngs/std.rs): seed bytes
    [1,0,0,0, 23,0,0,0, 200,1,0,0, 210,30,0,0, 0...] give next_u64() == 10719222850664546238, and the
    generator built from that stream (`from_rng`) gives 14064965282130556830.  That fixes the round count
    (12), the word order of next_u64 and the 64-bit block counter starting at 0;
  * PCG32 XSH-RR output function — the reference demo of pcg-c-basic (seed 42, sequence 54):
    a15c02b7 7b47f409 ba1d3330 83d2f293 bfa4784b cbed606e.  `seed_from_u64` (rand_core) uses the same output
    function on the post-advance state with its own increment, checked here against that formula.

What stays unpinned: nothing in this container can run `WeightedIndex` / `SliceRandom::shuffle` of the real
crate, so those two follow the crate's published algorithm (c4_rng.cuh) and are pinned oracle <-> product only.
"""

import ctypes as C
import struct

import numpy as np

import oracle

CHACHA12_ZERO = ("9bf49a6a0755f953811fce125f2683d50429c3bb49e074147e0089a52eae155f"
                 "0564f879d27ae3c02ce82834acfa8c793a629f2ca0de6919610be82f411326be")
CHACHA20_ZERO = ("76b8e0ada0f13d90405d6ae55386bd28bdd219b8a08ded1aa836efcc8b770dc7"
                 "da41597c5157488d7724e03fb8d84a376a43b8f41518a11cc387b669b2ee6586")
CHACHA8_ZERO = ("3e00ef2f895f40d67f5bb8e81f09a5a12c840ec3ce9a7f3b181be188ef711a1e"
                "984ce172b9216f419f445367456d5619314a42a3da86b001387bfdb80e0cfe42")
STDRNG_SEED_WORDS = [1, 23, 456, 7890, 0, 0, 0, 0]  # the test's seed bytes as little-endian u32 words
STDRNG_X0, STDRNG_X1 = 10719222850664546238, 14064965282130556830
PCG_DEMO = [0xA15C02B7, 0x7B47F409, 0xBA1D3330, 0x83D2F293, 0xBFA4784B, 0xCBED606E]
MUL, MASK = 6364136223846793005, (1 << 64) - 1


def _oracle_block(key, counter, rounds):
    out = (C.c_uint32 * 16)()
    oracle.lib().c4o_chacha_block((C.c_uint32 * 8)(*key), counter, rounds, out)
    return list(out)


def _hex(words):
    return b"".join(struct.pack("<I", w) for w in words).hex()


def _xsh_rr(state):
    xs = (((state >> 18) ^ state) >> 27) & 0xFFFFFFFF
    rot = state >> 59
    return ((xs >> rot) | (xs << ((32 - rot) & 31))) & 0xFFFFFFFF


def test_oracle_chacha_matches_published_keystreams():
    assert _hex(_oracle_block([0] * 8, 0, 12)) == CHACHA12_ZERO
    assert _hex(_oracle_block([0] * 8, 0, 20)) == CHACHA20_ZERO
    assert _hex(_oracle_block([0] * 8, 0, 8)) == CHACHA8_ZERO


def test_oracle_reproduces_rand_stdrng_value_stability_test():
    b = _oracle_block(STDRNG_SEED_WORDS, 0, 12)
    assert b[0] | (b[1] << 32) == STDRNG_X0
    b2 = _oracle_block(b[2:10], 0, 12)  # from_rng: the next 32 bytes of the stream seed a new generator
    assert b2[0] | (b2[1] << 32) == STDRNG_X1


def test_pcg32_output_function_matches_the_reference_demo():
    inc = (54 << 1) | 1
    state = (0 * MUL + inc) & MASK
    state = (state + 42) & MASK
    state = (state * MUL + inc) & MASK
    outs = []
    for _ in range(6):
        old = state
        state = (old * MUL + inc) & MASK
        outs.append(_xsh_rr(old))
    assert outs == PCG_DEMO


def _seed_from_u64(seed):
    """rand_core's SeedableRng::seed_from_u64: advance first, output from the NEW state, fixed increment."""
    inc, state, key = 11634580027462260723, seed, []
    for _ in range(8):
        state = (state * MUL + inc) & MASK
        key.append(_xsh_rr(state))
    return key


def test_product_stdrng_matches_the_same_vectors():
    from c4a0_b200 import _lib as L

    lib = L.lib()

    def words(key, n):
        k = np.array(key, np.uint32)
        out = np.zeros(n, np.uint32)
        lib.c4a0_host_stdrng_words(L.ptr(k), L.ptr(out), n)
        return out.tolist()

    assert _hex(words([0] * 8, 16)) == CHACHA12_ZERO
    w = words(STDRNG_SEED_WORDS, 10)
    assert w[0] | (w[1] << 32) == STDRNG_X0
    w2 = words(w[2:10], 2)
    assert w2[0] | (w2[1] << 32) == STDRNG_X1
    # 40 words cross two block boundaries: the counter increments by one per block
    long = words(STDRNG_SEED_WORDS, 40)
    assert long[16:32] == _oracle_block(STDRNG_SEED_WORDS, 1, 12) and long[32:40] == _oracle_block(STDRNG_SEED_WORDS, 2, 12)[:8]
    for seed in (0, 1, 42, 2**63 + 12345, 2**64 - 1, 7 * (42 + 5)):
        key = np.zeros(8, np.uint32)
        lib.c4a0_host_seed_to_key(seed, L.ptr(key))
        assert key.tolist() == _seed_from_u64(seed)


def test_move_sampling_uses_the_first_word_of_the_seeded_stream():
    """mcts.rs:216-220: one draw from StdRng::seed_from_u64(game_id * (42 + n_moves)).  With weights that put
    all mass on one column the draw is irrelevant; with two equal weights the column is decided by the top
    23 bits of word 0 of the seeded stream — recomputed here from the pinned pieces."""
    from c4a0_b200 import engine as E

    for seed in (0, 3, 99, 2**40 + 7):
        b = _oracle_block(_seed_from_u64(seed), 0, 12)
        v01 = np.float32(np.uint32((b[0] >> 9) | 0x3F800000).view(np.float32) - np.float32(1.0))
        want = 2 if v01 < np.float32(0.5) else 5
        w = [0, 0, 0.5, 0, 0, 0.5, 0]
        assert oracle.weighted_index_sample(w, seed) == want
        assert E.host_sample(w, 1.0, seed)[1] == want
