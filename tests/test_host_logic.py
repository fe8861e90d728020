"""CPU-only tests of the host side: the Board value type against the reference's goldens, the BN-folded weight blob,
record <-> GameData plumbing, and that libc4b200.so loads and exports every symbol include/c4b200.h declares."""
import ctypes
import json
import os
import re

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden

from connect4_b200.board import Board, make_random_ips
from connect4_b200.utils import Result, RESULT_FROM_CODE


def test_board_result_kats():
    """reference tests/board_test.py:152-161"""
    kat = json.load(open(os.path.join(GOLDEN, "board_kat.json")))
    for case in kat["result_cases"]:
        b = Board.from_pieces(o_pieces=np.array(case["o"], np.bool_), x_pieces=np.array(case["x"], np.bool_))
        assert b.result == (Result(case["ans"]) if case["ans"] is not None else None)
        assert (int(b.color[0]), int(b.color[1]), b.age) == (case["c0"], case["c1"], case["age"])
        assert [int(h) for h in b.height] == case["height"]


def test_board_valid_moves_kats():
    """reference tests/board_test.py:164-247"""
    kat = json.load(open(os.path.join(GOLDEN, "board_kat.json")))
    for case in kat["valid_move_cases"]:
        b = Board.from_pieces(o_pieces=np.array(case["o"], np.bool_), x_pieces=np.array(case["x"], np.bool_))
        assert b.valid_moves == set(case["valid"])


def test_board_playouts():
    g = golden("board_playouts.npz")
    n = len(g["c0"])
    b = None
    for i in range(n):
        if i == 0 or g["game"][i] != g["game"][i - 1]:
            b = Board()
        assert (int(b.color[0]), int(b.color[1]), b.age) == (int(g["c0"][i]), int(g["c1"][i]), int(g["age"][i]))
        assert b.result == RESULT_FROM_CODE[int(g["result"][i])]
        assert sum(1 << m for m in b.valid_moves) == int(g["valid"][i])
        f = b.create_fliplr()
        assert (int(f.color[0]), int(f.color[1])) == (int(g["f0"][i]), int(g["f1"][i]))
        assert b.symmetrical == bool(g["sym"][i])
        assert (np.packbits(b.to_array().reshape(-1)) == g["planes"][i]).all()
        rb = Board.from_bitboards(b.color[0], b.color[1])
        assert rb == b and rb.age == b.age and rb.result == b.result and (rb.height == b.height).all()
        if i + 1 < n and g["game"][i + 1] == g["game"][i]:
            diff = (int(g["c0"][i + 1]) ^ int(b.color[0])) | (int(g["c1"][i + 1]) ^ int(b.color[1]))
            b.make_move((diff.bit_length() - 1) // 7)


def test_make_random_ips():
    kat = json.load(open(os.path.join(GOLDEN, "board_kat.json")))
    assert [len(make_random_ips(p)) for p in range(4)] == kat["make_random_ips_counts"]
    assert sorted([int(b.color[0]), int(b.color[1])] for b in make_random_ips(2)) == kat["make_random_ips_2"]


def _blob_forward(blob, planes):
    """fp32 forward straight from the folded blob (torch CPU) -- checks the folding algebra, not the CUDA kernel"""
    import torch
    import torch.nn.functional as F
    t = torch.as_tensor(blob)
    Fi, R = int(blob[1]), int(blob[2])
    assert int(blob[3]) >> 8 == 0                   # default operand dtype fp16, kernel auto
    pos = [4]

    def take(*shape):
        n = int(np.prod(shape))
        v = t[pos[0]:pos[0] + n].reshape(*shape)
        pos[0] += n
        return v
    lk = lambda z: F.leaky_relu(z, 0.01)
    x = torch.as_tensor(planes, dtype=torch.float32)
    h = lk(F.conv2d(x, take(Fi, 3, 3, 3), take(Fi), padding=1))
    for _ in range(R):
        o = lk(F.conv2d(h, take(Fi, Fi, 3, 3), take(Fi), padding=1))
        h = lk(F.conv2d(o, take(Fi, Fi, 3, 3), take(Fi), padding=1) + h)
    v = lk(F.conv2d(h, take(1, Fi, 1, 1), take(1))).reshape(len(x), -1)
    v = lk(F.linear(v, take(42, 42), take(42)))
    v = torch.tanh(F.linear(v, take(1, 42), take(1)))
    v = ((v + take(1)) * take(1)).reshape(-1)
    p = lk(F.conv2d(h, take(2, Fi, 1, 1), take(2))).reshape(len(x), -1)
    p = torch.softmax(F.linear(p, take(7, 84), take(7)), dim=1)
    assert pos[0] == len(blob)
    return v.numpy(), p.numpy()


def test_fold_state_dict_matches_reference_outputs():
    import torch
    from oracle import net_ref as nr
    from connect4_b200.neural.weights import fold_state_dict, net_shape
    torch.set_num_threads(2)
    g = golden("net_outputs.npz")
    sd = nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz"))
    assert net_shape(sd) == (32, 3, 4)
    blob = fold_state_dict(sd)
    planes = nr.planes_from_bitboards(g["c0"][:512], g["c1"][:512])
    v, p = _blob_forward(blob, planes)
    assert np.abs(v - g["value"][:512]).max() < 2e-5
    assert np.abs(p - g["prior"][:512]).max() < 2e-5
    big = nr.random_state(0, 64, 6, 6)
    assert net_shape(big) == (64, 6, 6)
    v, p = _blob_forward(fold_state_dict(big), planes[:256])
    assert np.abs(v - g["big_value"]).max() < 2e-5 and np.abs(p - g["big_prior"]).max() < 2e-5


def test_init_state_dict_matches_reference_init():
    """ModelWrapper's random init draws the same parameters as the reference's Net under the same torch seed"""
    import torch
    from oracle import net_ref as nr
    from connect4_b200.neural.config import NetConfig
    from connect4_b200.neural.model import _init_state_dict
    torch.manual_seed(0)
    mine = _init_state_dict(NetConfig())
    ref = nr.random_state(0)
    assert set(mine) == set(ref)
    for k in ref:
        assert np.array_equal(mine[k].numpy(), ref[k]), k


def test_records_to_gamedata_and_back():
    from connect4_b200.engine import RECORD_DTYPE
    from connect4_b200.neural.training_game import games_from_records
    rec = np.zeros(5, RECORD_DTYPE)
    b = Board()
    moves = [3, 3, 2]
    for i, m in enumerate(moves):
        rec[i]["c0"], rec[i]["c1"] = int(b.color[0]), int(b.color[1])
        rec[i]["move"], rec[i]["ply"], rec[i]["game_id"], rec[i]["result"] = m, i, 7, 2
        rec[i]["policy"] = np.arange(7) / 21.0
        rec[i]["search_value"] = np.nan if i == 0 else 0.25
        b.make_move(m)
    rec[3]["game_id"], rec[3]["ply"], rec[3]["move"], rec[3]["result"] = 2, 0, 1, 1
    rec[4]["game_id"], rec[4]["ply"], rec[4]["move"], rec[4]["result"], rec[4]["c0"] = 2, 1, 0, 1, 1 << 7
    games = games_from_records(rec[[4, 0, 3, 2, 1]])
    assert [g.moves for g in games] == [[1, 0], [3, 3, 2]]
    assert games[1].result == Result.o_win and games[0].result == Result.draw
    assert games[1].values[0] is None and games[1].values[1] == 0.25
    assert games[1].boards[2].age == 2 and games[1].data.values == [1.0, 1.0, 1.0]


def test_shared_library_exports_every_declared_symbol():
    from connect4_b200 import _build, _lib
    header = open(os.path.join(ROOT, "include", "c4b200.h")).read()
    declared = set(re.findall(r"\b(c4_[a-z0-9_]+)\s*\(", header))
    declared -= {"c4_ctx", "c4_net"}
    assert len(declared) >= 30
    path = _build.build()
    L = ctypes.CDLL(path)
    for name in sorted(declared):
        assert hasattr(L, name), "missing export: " + name
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    assert _lib.load().c4_abi_version() == 1
    assert ctypes.sizeof(_lib.RecordC) == 64


def test_no_cpu_fallback():
    """the product path refuses to run without a GPU instead of silently computing on the host"""
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from connect4_b200 import _lib
    from connect4_b200.evaluators import Evaluator, evaluate_centre_with_prior
    from connect4_b200.mcts import MCTS, MCTSConfig
    with pytest.raises(_lib.C4Error):
        MCTS("t", MCTSConfig(8), Evaluator(evaluate_centre_with_prior)).make_move(Board())
    with pytest.raises(_lib.C4Error):
        from connect4_b200.neural.model import ModelWrapper
        ModelWrapper()


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "connect4_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("by the oracle", ""), os.path.join(dirpath, f)


def test_match_chooses_the_lock_step_path_only_for_deterministic_device_players():
    """Match.play_batched is taken for MCTS players with device evaluators and no host randomness; everything else keeps
    the reference's one-game-at-a-time path (match.py:42-50)"""
    from functools import partial
    from connect4_b200 import evaluators as evl
    from connect4_b200.match import Match, _batchable
    from connect4_b200.mcts import MCTS, MCTSConfig, device_kind
    from connect4_b200.player import HumanPlayer

    class FakeModel():                       # anything owning a c4_net handle is a network evaluator
        c4_net = object()

    centre = MCTS("c", MCTSConfig(50), evl.Evaluator(evl.evaluate_centre_with_prior))
    net = MCTS("n", MCTSConfig(50), evl.Evaluator(partial(evl.evaluate_nn, model=FakeModel())))
    bare = MCTS("b", MCTSConfig(50), FakeModel())
    host = MCTS("h", MCTSConfig(50), evl.Evaluator(lambda board: (0.5, np.ones(7) / 7)))
    noisy = MCTS("z", MCTSConfig(50, 19652, 1.25, 0.3, 0.25, 0), evl.Evaluator(evl.evaluate_centre_with_prior))
    sampled = MCTS("s", MCTSConfig(50, num_sampling_moves=6), evl.Evaluator(evl.evaluate_centre_with_prior))
    assert [device_kind(p.evaluator)[0] for p in (centre, net, bare, host)] == ["centre", "net", "net", "external"]
    assert [_batchable(p) for p in (centre, net, bare, host, noisy, sampled, HumanPlayer("me"))] == \
        [True, True, True, False, False, False, False]
    m = Match(False, centre, net, plies=1, switch=True)          # 7 one-ply openings, each played twice
    assert m.n == 7 and len(m.games) == 14
    assert all(g._player_o.evaluator is centre.evaluator for g in m.games[:7])
    assert all(g._player_o.evaluator is net.evaluator for g in m.games[7:])
    assert all(g.player_to_move() is g._player_x for g in m.games)      # after one ply x is to move


def test_adaptive_tower_count_decisions():
    """the split engine's controller (csrc/c4_split.cu, sp_adapt_next) on load signals measured on a B200 (profiles/README.md,
    time profile of a generation): host arithmetic only, reached through a test hook of the library"""
    import ctypes
    from connect4_b200 import _build
    L = ctypes.CDLL(_build.build())
    f = L.c4_split_adapt_next
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int] * 3 + [ctypes.c_ulonglong] * 4

    def nxt(n_net, boards_per_strip, idle_share, games=4096, sms=148):
        strips = 10000
        return f(sms, games, n_net, strips, int(round(boards_per_strip * strips)), int(round((1.0 - idle_share) * 1e9)), int(round(idle_share * 1e9)))

    # middle game: saturated towers (15+ boards per strip) hold the trees back -> more towers, at most 16 at a time
    assert nxt(56, 15.33, 0.34) == 72
    assert nxt(64, 15.6, 0.43) == 80
    assert nxt(72, 15.5, 0.25) in (80, 84)
    # balanced: towers full, tree warps idle ~5 % -> stays within one step of 4
    assert abs(nxt(80, 15.1, 0.062) - 80) <= 4
    # tail: short strips, trees never wait -> towers become tree SMs
    assert nxt(72, 5.0, 0.0) == 56
    assert nxt(80, 3.7, 0.0) == 64                                   # (wants ~54: one step of 16)
    assert nxt(56, 12.6, 0.027) == 56
    # cold start: the trees wait although the strips are short -- latency, not capacity: no tower is taken away
    assert nxt(56, 4.18, 0.10) == 56
    assert nxt(56, 1.0, 0.069) == 56
    # bounds: 40..96 of 148 SMs, and a tree CTA owns at most 256 game slots (16,384 games -> at least 64 tree CTAs)
    assert nxt(96, 16.0, 0.6) == 96
    assert nxt(40, 1.0, 0.0) == 40
    assert nxt(84, 16.0, 0.5, games=16384) == 84
    assert nxt(66, 16.0, 0.5, sms=132) <= (132 * 96 + 74) // 148     # a smaller device scales the bounds
