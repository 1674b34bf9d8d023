"""Host-side behaviour of the drop-in `c4a0_rust` classes (no GPU): result classes, pickling / CBOR,
split_train_test, player0_score — checked against the oracle's restatement of types.rs / pybridge.rs
and against an independent CBOR encoder."""

import pickle

import numpy as np
import pytest

import c4a0_rust as R
import oracle
from c4a0_rust import _cbor


def _result_from_oracle(n_games=9, n_iter=12):
    reqs = [(10 + i, i % 2, (i + 1) % 2) for i in range(n_games)]
    out = oracle.self_play(reqs, 4, n_iter, 2.0, 0.01, evaluator="hash")
    games = []
    for r, ss in zip(reqs, out.samples):
        samples = [R.Sample(s.pos.mask, s.pos.value, list(s.policy), s.q_penalty, s.q_no_penalty) for s in ss]
        games.append(R.GameResult(R.GameMetadata(*r), samples))
    return R.PlayGamesResult._from_results(games), out


def test_module_surface():
    assert (R.N_COLS, R.N_ROWS, R.BUF_N_CHANNELS) == (7, 6, 2)  # lib.rs:28-30
    for name in ("GameMetadata", "GameResult", "Sample", "PlayGamesResult", "play_games", "run_tui"):
        assert hasattr(R, name)
    assert R.PlayGamesResult.__module__ == "c4a0_rust"  # pybridge.rs:57-59
    md = R.GameMetadata(3, 1, 2)
    assert (md.game_id, md.player0_id, md.player1_id) == (3, 1, 2)
    with pytest.raises(TypeError):
        R.GameMetadata()  # types.rs:52-59: all three are required
    with pytest.raises(AttributeError):
        md.game_id = 4
    with pytest.raises(OverflowError):
        R.GameMetadata(-1, 0, 0)
    assert len(R.PlayGamesResult().results) == 0


def test_sample_views():
    res, out = _result_from_oracle()
    for g, ss in zip(res.results, out.samples):
        for s, o in zip(g.samples, ss):
            pos, pol, qp, qn = s.to_numpy()
            assert pos.dtype == np.float32 and pos.shape == (2, 6, 7)
            assert np.array_equal(pos, oracle.planes(o.pos))
            assert pol.tobytes() == np.array(list(o.policy), np.float32).tobytes()
            assert qp.shape == () and float(qp) == o.q_penalty and float(qn) == o.q_no_penalty
            assert s.pos_str() == oracle.to_str(o.pos)
            f = s.flip_h()
            fo = oracle.lib().c4o_flip_h(o.pos)
            assert (f._mask, f._value) == (fo.mask, fo.value)
            assert np.array_equal(f.to_numpy()[1], pol[::-1])
            assert f.flip_h() == s
        assert g.player0_score() == oracle.player0_score(ss)
    with pytest.raises(RuntimeError):
        R.GameResult(R.GameMetadata(0, 0, 0), res.results[0].samples[:1]).player0_score()


def test_pickle_and_cbor_round_trip():
    res, _ = _result_from_oracle()
    blob = res.to_cbor()
    again = R.PlayGamesResult.from_cbor(blob)
    assert [g.samples for g in again.results] == [g.samples for g in res.results]
    assert [g.metadata for g in again.results] == [g.metadata for g in res.results]
    assert again.to_cbor() == blob
    p = pickle.loads(pickle.dumps(res))
    assert p.to_cbor() == blob
    assert b"c4a0_rust" in pickle.dumps(res) and b"_native" not in pickle.dumps(res)
    both = res + again
    assert len(both.results) == 2 * len(res.results)
    assert both.unique_positions() == res.unique_positions() > 0
    for bad in (b"", b"\x01", b"\xa1\x67results\x01", blob[:-3]):
        with pytest.raises(ValueError):
            R.PlayGamesResult.from_cbor(bad)


def test_cbor_bytes_follow_serde_cbor_conventions():
    """Struct -> map with field-name keys in declaration order; u64 shortest form; f32 as half when
    lossless (serde_cbor 0.11 `serialize_f32`).  Cross-checked with cbor2's canonical float rule."""
    cbor2 = pytest.importorskip("cbor2")
    s = R.Sample(0x0810204081, 1, [0.5, 0.25, 0.0, 1 / 7, 0, 0, 0], -1.0, 0.97)
    res = R.PlayGamesResult._from_results([R.GameResult(R.GameMetadata(2**40, 0, 300), [s])])
    doc = cbor2.loads(res.to_cbor())
    g = doc["results"][0]
    assert list(doc) == ["results"] and list(g) == ["metadata", "samples"]
    assert list(g["metadata"]) == ["game_id", "player0_id", "player1_id"]
    assert list(g["samples"][0]) == ["pos", "policy", "q_penalty", "q_no_penalty"]
    assert list(g["samples"][0]["pos"]) == ["mask", "value"]
    assert g["metadata"]["game_id"] == 2**40 and g["samples"][0]["pos"]["mask"] == 0x0810204081
    for v in (0.5, -1.0, 0.0, float(np.float32(1 / 7)), float(np.float32(0.97)), 65504.0, 1e-8, float("inf")):
        assert _cbor.encode_f32(v) == cbor2.dumps(float(np.float32(v)), canonical=True), v
    assert len(_cbor.encode_f32(0.5)) == 3 and len(_cbor.encode_f32(np.float32(1 / 7))) == 5


def test_split_train_test_matches_reference_semantics():
    res, _ = _result_from_oracle(n_games=11)
    games = res.results
    for frac, seed in ((0.8, 1337), (0.5, 0), (0.0, 5), (1.0, 5), (0.45, 2**63)):
        train, test = res.split_train_test(frac, seed)
        order = oracle.shuffle_indices(seed, len(games))  # StdRng shuffle restated independently
        n_train = int(np.floor(np.float32(len(games)) * np.float32(frac) + np.float32(0.5)))
        exp_train = [s for i in order[:n_train] for s in games[i].samples]
        exp_test = [s for i in order[n_train:] for s in games[i].samples]
        assert train == exp_train and test == exp_test
        assert res.split_train_test(frac, seed) == (train, test)  # deterministic, non-mutating
    assert [g.samples for g in res.results] == [g.samples for g in games]


def test_to_arrays_bulk_export():
    res, out = _result_from_oracle()
    pos, pol, qp, qn = res.to_arrays()
    flat = [s for ss in out.samples for s in ss]
    assert pos.shape == (len(flat), 2, 6, 7) and pol.shape == (len(flat), 7)
    for i in (0, len(flat) // 2, len(flat) - 1):
        assert np.array_equal(pos[i], oracle.planes(flat[i].pos))
        assert qp[i] == np.float32(flat[i].q_penalty) and qn[i] == np.float32(flat[i].q_no_penalty)


def test_unsupported_entry_points_say_so():
    with pytest.raises(NotImplementedError):
        R.run_tui(lambda m, p: None, 10, 1.0, 0.01)
    with pytest.raises(NotImplementedError):
        R.PlayGamesResult().score_policies("a", "b", "c")


def _python_cbor(res):
    """PlayGamesResult -> bytes with the Python codec (c4a0_rust/_cbor.py): the cross-check of csrc/cbor.cu."""
    F = _cbor.F32
    games = []
    for g in res.results:
        md = g.metadata
        games.append({
            "metadata": {"game_id": md.game_id, "player0_id": md.player0_id, "player1_id": md.player1_id},
            "samples": [{"pos": {"mask": s._mask, "value": s._value}, "policy": [F(x) for x in s._policy.tolist()],
                         "q_penalty": F(s._q_penalty), "q_no_penalty": F(s._q_no_penalty)} for s in g.samples],
        })
    return _cbor.dumps({"results": games})


def test_bulk_cbor_codec_equals_the_python_codec_byte_for_byte():
    res, _ = _result_from_oracle(n_games=9)
    assert res.to_cbor() == _python_cbor(res)
    # float edge cases: half normals / subnormals / limits, values that need a single, infinities, -0, NaN;
    # integers at every head width
    vals = [0.0, -0.0, 1.0, -1.0, 0.5, 65504.0, 65505.0, 65520.0, 2.0**-14, 2.0**-15, 2.0**-24, 2.0**-25, 3 * 2.0**-24,
            1023 * 2.0**-24, 1025 * 2.0**-25, 1e-8, 1 / 7, 0.97, 1e30, float("inf"), float("-inf"), float("nan"), 6.1035156e-05]
    ints = [0, 23, 24, 255, 256, 65535, 65536, 2**32 - 1, 2**32, 2**64 - 1]
    samples = []
    for i in range(0, len(vals) - 8, 3):
        samples.append(R.Sample(ints[i % len(ints)], ints[(i + 3) % len(ints)], vals[i : i + 7], vals[i + 7], vals[(i + 8) % len(vals)]))
    weird = R.PlayGamesResult._from_results([R.GameResult(R.GameMetadata(ints[k], ints[-1 - k], ints[(2 * k) % len(ints)]), samples)
                                             for k in range(len(ints))])
    blob = weird.to_cbor()
    assert blob == _python_cbor(weird)
    back = R.PlayGamesResult.from_cbor(blob)
    assert np.array_equal(back._meta, weird._meta)
    for f in ("n_samples", "mask", "value"):
        assert np.array_equal(getattr(back._soa, f), getattr(weird._soa, f)), f
    for f in ("policy", "q_penalty", "q_no_penalty"):  # bit patterns, NaN payload aside
        a, b = getattr(back._soa, f), getattr(weird._soa, f)
        same = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        assert same.all(), f
    # another encoder's bytes for the same document decode too (cbor2 writes canonical floats, other key widths)
    cbor2 = pytest.importorskip("cbor2")
    doc = cbor2.loads(res.to_cbor())
    doc["results"][0]["extra_field"] = [1, {"x": b"bytes"}]
    again = R.PlayGamesResult.from_cbor(cbor2.dumps(doc))
    assert again.to_cbor() == res.to_cbor()
    empty = R.PlayGamesResult()
    assert R.PlayGamesResult.from_cbor(empty.to_cbor()).to_cbor() == empty.to_cbor() == _cbor.dumps({"results": []})


def test_cbor_bytes_of_a_hand_encoded_document():
    """A document small enough to encode by hand from RFC 8949 and serde's data model (structs = maps keyed by
    field name in declaration order: pybridge.rs:60-63, types.rs:36-48, 65-71, 105-110, c4r.rs:14-17; integers with the
    shortest head; f32 as a half float where that is lossless).  No codec is involved in producing the expectation."""
    from c4a0_b200.engine import GameSamples

    def text(s):
        assert len(s) < 24
        return bytes([0x60 | len(s)]) + s.encode()

    half = {0.5: b"\xf9\x38\x00", -1.0: b"\xf9\xbc\x00", 0.25: b"\xf9\x34\x00", 0.0: b"\xf9\x00\x00"}
    third = b"\xfa" + np.float32(1 / 3).tobytes()[::-1]  # not representable as a half: stays a single
    want = (
        b"\xa1" + text("results") + b"\x81"                                   # {results: [ one game
        + b"\xa2" + text("metadata") + b"\xa3" + text("game_id") + b"\x18\x2a" + text("player0_id") + b"\x00"
        + text("player1_id") + b"\x1b" + (2 ** 40).to_bytes(8, "big")          # u64 that needs the 8-byte head
        + text("samples") + b"\x82"                                            # two samples
        + b"\xa4" + text("pos") + b"\xa2" + text("mask") + b"\x01" + text("value") + b"\x00"
        + text("policy") + b"\x87" + half[0.5] * 6 + third
        + text("q_penalty") + half[-1.0] + text("q_no_penalty") + half[0.25]
        + b"\xa4" + text("pos") + b"\xa2" + text("mask") + b"\x19\x01\x81" + text("value") + b"\x18\x80"
        + text("policy") + b"\x87" + half[0.0] * 7
        + text("q_penalty") + half[0.0] + text("q_no_penalty") + half[0.5]
    )
    M = 43
    soa = GameSamples(np.array([2], np.uint32), np.zeros((1, M), np.uint64), np.zeros((1, M), np.uint64),
                      np.zeros((1, M, 7), np.float32), np.zeros((1, M), np.float32), np.zeros((1, M), np.float32))
    soa.mask[0, :2] = [1, 0x181]
    soa.value[0, :2] = [0, 0x80]
    soa.policy[0, 0] = [0.5] * 6 + [np.float32(1 / 3)]
    soa.q_penalty[0, :2] = [-1.0, 0.0]
    soa.q_no_penalty[0, :2] = [0.25, 0.5]
    res = R.PlayGamesResult._from_soa(np.array([[42, 0, 2 ** 40]], np.uint64), soa)
    assert res.to_cbor() == want
    back = R.PlayGamesResult.from_cbor(want)
    assert back.results[0].metadata.player1_id == 2 ** 40 and len(back.results[0].samples) == 2
    assert back.to_cbor() == want
