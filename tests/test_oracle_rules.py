"""Pins the CPU oracle's rules against every known-answer test the reference holds for
rust/src/c4r.rs (tests at c4r.rs:474-653 and rust/proptest-regressions/c4r.txt)."""

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import oracle as O

DRAW_MOVES = (
    [0, 1, 2, 3, 4, 5] * 3 + [5, 4, 3, 2, 1, 0] * 3 + [6] * 6
)  # c4r.rs:505-516

cols_strategy = st.lists(st.integers(0, 6), min_size=0, max_size=500)  # c4r.rs:614


def test_playing_moves_works():  # c4r.rs:474-487
    pos = O.Pos(0, 0)
    for col in range(7):
        for row in range(6):
            pos = O.make_move(pos, col)
            assert O.lib().c4o_get(pos, row, col) == 0  # Opponent after inversion
        assert not O.legal_moves(pos)[col]
        assert O.make_move(pos, col) is None


def test_row_win():  # c4r.rs:489-495
    assert O.terminal_state(O.from_moves([0, 0, 1, 1, 2, 2, 3])) == O.OPPONENT_WIN


def test_col_win():  # c4r.rs:497-501
    assert O.terminal_state(O.from_moves([6, 0, 6, 0, 6, 0, 6])) == O.OPPONENT_WIN


def test_draw():  # c4r.rs:503-520
    assert O.terminal_state(O.from_moves(DRAW_MOVES)) == O.DRAW


def test_to_str():  # c4r.rs:522-546
    pos = O.from_moves(DRAW_MOVES)
    expected = "\n".join(
        [
            "🔵🔴🔵🔴🔵🔴🔵",
            "🔵🔴🔵🔴🔵🔴🔴",
            "🔵🔴🔵🔴🔵🔴🔵",
            "🔴🔵🔴🔵🔴🔵🔴",
            "🔴🔵🔴🔵🔴🔵🔵",
            "🔴🔵🔴🔵🔴🔵🔴",
        ]
    )
    assert O.to_str(pos) == expected
    assert O.from_str(expected).key() == pos.key()


def _legal_str(pos):
    return "".join("O" if l else "X" for l in O.legal_moves(pos))


def test_legal_moves():  # c4r.rs:548-579
    pos = O.Pos(0, 0)
    assert _legal_str(pos) == "OOOOOOO"
    moves = [0, 1, 2, 3, 4, 5] * 3 + [5, 4, 3, 2, 1, 0] * 2
    pos = O.from_moves(moves)
    assert _legal_str(pos) == "OOOOOOO"
    for col, expect in zip([5, 4, 3, 2, 1, 0], ["OOOOOXO", "OOOOXXO", "OOOXXXO", "OOXXXXO", "OXXXXXO", "XXXXXXO"]):
        pos = O.make_move(pos, col)
        assert _legal_str(pos) == expect
    for _ in range(6):
        pos = O.make_move(pos, 6)
    assert _legal_str(pos) == "XXXXXXX"


def test_flip_h_symmetrical():  # c4r.rs:602-608
    pos = O.from_moves([3, 3, 3])
    flipped = O.lib().c4o_flip_h(pos)
    assert pos.key() == flipped.key()
    assert pos.key() == O.lib().c4o_flip_h(flipped).key()


# rust/proptest-regressions/c4r.txt:7-8 — the reference's only literal 64-bit mask/value vectors
GOLDEN_BOARDS = [
    (
        "🔵⚫⚫⚫⚫⚫⚫ 🔴⚫⚫⚫⚫⚫⚫ 🔵⚫⚫⚫⚫⚫⚫ 🔴⚫⚫⚫⚫⚫⚫ 🔵⚫⚫⚫⚫⚫⚫ 🔴⚫🔵⚫⚫⚫⚫",
        "0000000000000000000000000000100000010000001000000100000010000101",
        "0000000000000000000000000000000000010000000000000100000000000001",
    ),
    (
        "🔵⚫⚫⚫⚫⚫⚫ 🔴⚫⚫⚫⚫⚫⚫ 🔵⚫⚫⚫⚫⚫⚫ 🔴⚫⚫⚫⚫⚫🔵 🔵🔴⚫⚫⚫⚫🔵 🔴🔵🔴⚫⚫⚫🔴",
        "0000000000000000000000000000100000010000001100000110000111000111",
        "0000000000000000000000000000000000010000000000000100000101000101",
    ),
]


@pytest.mark.parametrize("board,mask,value", GOLDEN_BOARDS)
def test_golden_bit_layout(board, mask, value):
    s = board.replace(" ", "\n")
    pos = O.from_str(s)
    assert pos.mask == int(mask, 2)
    assert pos.value == int(value, 2)
    assert O.to_str(O.Pos(int(mask, 2), int(value, 2))) == s


def test_win_masks_count_and_shape():  # c4r.rs:165-224
    wm = O.win_masks()
    assert len(wm) == 69 and len(set(wm)) == 69
    assert all(bin(m).count("1") == 4 and m < (1 << 42) for m in wm)
    # first horizontal, first vertical, first of each diagonal family
    assert wm[0] == 0b1111
    assert wm[24] == (1 | 1 << 7 | 1 << 14 | 1 << 21)
    assert wm[45] == (1 | 1 << 8 | 1 << 16 | 1 << 24)
    assert wm[57] == (1 << 21 | 1 << 15 | 1 << 9 | 1 << 3)


@settings(max_examples=300, deadline=None)
@given(cols_strategy)
def test_prop_flip_h_involution(cols):  # c4r.rs:632-637
    pos = O.random_pos(cols)
    f = O.lib().c4o_flip_h(pos)
    assert O.lib().c4o_flip_h(f).key() == pos.key()


@settings(max_examples=300, deadline=None)
@given(cols_strategy)
def test_prop_string_round_trip(cols):  # c4r.rs:639-644
    pos = O.random_pos(cols)
    assert O.from_str(O.to_str(pos)).key() == pos.key()


@settings(max_examples=200, deadline=None)
@given(cols_strategy)
def test_prop_planes_match_bits(cols):  # c4r.rs:378-392
    pos = O.random_pos(cols)
    pl = O.planes(pos).reshape(2, 42)
    me = pos.mask & pos.value
    opp = pos.mask & ~pos.value
    for b in range(42):
        assert pl[0, b] == (me >> b & 1)
        assert pl[1, b] == (opp >> b & 1)


def test_terminal_value():  # c4r.rs:253-263
    pos = O.from_moves([0, 0, 1, 1, 2, 2, 3])
    qp, qn = O.terminal_value(pos, 0.01)
    assert qn == -1.0
    assert qp == np.float32(-1.0) + np.float32(0.01) * np.float32(7)
    assert O.terminal_value(O.Pos(0, 0), 0.01) is None
    assert O.terminal_value(O.from_moves(DRAW_MOVES), 0.01) == (0.0, 0.0)
