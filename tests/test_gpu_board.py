"""GPU bitboard engine (c4_board_* through the C ABI) vs the oracle and the reference's goldens. Bit-exact."""
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN, bits, golden, random_positions

pytestmark = pytest.mark.gpu


def test_board_ops_on_reference_playouts(oracle):
    from connect4_b200.board import BoardBatch
    g = golden("board_playouts.npz")
    bb = BoardBatch(g["c0"], g["c1"])
    res = bb.result().cpu().numpy()
    assert (res == g["result"]).all()
    assert (bb.legal_mask(bb.result()).cpu().numpy() == g["valid"]).all()
    running = bb.legal_mask().cpu().numpy()               # result=None: treat every game as running
    for i in range(0, len(res), 97):
        assert running[i] == oracle.legal_mask(int(g["c0"][i]), int(g["c1"][i]))
    f0, f1 = bb.fliplr().numpy()
    assert (f0 == g["f0"]).all() and (f1 == g["f1"]).all()
    planes = bb.to_planes("uint8").cpu().numpy()
    assert (np.packbits(planes.reshape(len(res), -1), axis=1) == g["planes"]).all()
    assert (bb.to_planes("float32").cpu().numpy() == planes.astype(np.float32)).all()
    back = BoardBatch.from_planes(planes[:, 1], planes[:, 2]).numpy()
    assert (back[0] == g["c0"]).all() and (back[1] == g["c1"]).all()
    assert (bb.has_win(0).cpu().numpy() == np.array([oracle.has_win(int(x)) for x in g["c0"]])).all()
    assert (bb.has_win(1).cpu().numpy() == np.array([oracle.has_win(int(x)) for x in g["c1"]])).all()
    v = bb.evaluate_centre().cpu().numpy()
    want = np.array([oracle.evaluate_centre(int(a), int(b)) for a, b in zip(g["c0"], g["c1"])])
    assert (bits(v) == bits(want)).all()


def test_drop_replays_reference_games():
    from connect4_b200.board import BoardBatch
    g = golden("board_playouts.npz")
    n = len(g["c0"])
    same = g["game"][1:] == g["game"][:-1]
    idx = np.flatnonzero(same)
    diff = (g["c0"][idx + 1] ^ g["c0"][idx]) | (g["c1"][idx + 1] ^ g["c1"][idx])
    cols = np.array([(int(d).bit_length() - 1) // 7 for d in diff], np.int8)
    moves = np.full(n, -1, np.int8)
    moves[idx] = cols
    bb = BoardBatch(g["c0"].copy(), g["c1"].copy())
    res = bb.drop(moves).cpu().numpy()
    c0, c1 = bb.numpy()
    assert (c0[idx] == g["c0"][idx + 1]).all() and (c1[idx] == g["c1"][idx + 1]).all()
    assert (res[idx] == g["result"][idx + 1]).all()
    rest = np.flatnonzero(moves < 0)
    assert (c0[rest] == g["c0"][rest]).all()                  # negative move = untouched


def test_reference_kats_on_device():
    """reference tests/board_test.py (results + valid moves) through from_planes / result / legal_mask"""
    from connect4_b200.board import BoardBatch
    kat = json.load(open(os.path.join(GOLDEN, "board_kat.json")))
    cases = kat["result_cases"]
    bb = BoardBatch.from_planes(np.array([c["o"] for c in cases], np.uint8), np.array([c["x"] for c in cases], np.uint8))
    want = [-1 if c["ans"] is None else int(c["ans"] * 2) for c in cases]
    assert bb.result().cpu().numpy().tolist() == want
    cases = kat["valid_move_cases"]
    bb = BoardBatch.from_planes(np.array([c["o"] for c in cases], np.uint8), np.array([c["x"] for c in cases], np.uint8))
    masks = bb.legal_mask(bb.result()).cpu().numpy()
    for m, c in zip(masks, cases):
        assert [k for k in range(7) if m >> k & 1] == c["valid"]


def test_ragged_and_empty_sizes(oracle):
    from connect4_b200.board import BoardBatch
    for n in (0, 1, 31, 33, 1000003):
        c0, c1 = random_positions(3, min(n, 257))
        reps = (n + len(c0) - 1) // max(len(c0), 1) if n else 0
        a = np.tile(c0, reps)[:n] if n else np.zeros(0, np.uint64)
        b = np.tile(c1, reps)[:n] if n else np.zeros(0, np.uint64)
        bb = BoardBatch(a, b)
        r = bb.result().cpu().numpy()
        m = bb.legal_mask().cpu().numpy()
        assert r.shape == (n,) and (r == -1).all()
        for i in range(0, n, max(1, n // 50)):
            assert m[i] == oracle.legal_mask(int(a[i]), int(b[i]))
        f0, f1 = bb.fliplr().fliplr().numpy()                 # involution at the full size
        assert (f0 == a).all() and (f1 == b).all()
