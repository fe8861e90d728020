#!/usr/bin/env python3
"""Golden-vector generator: runs the UNMODIFIED reference (willis-richard/connect4, `oinkoink`) from
/root/reference and stores its outputs as small fixtures next to this script.

TEST INFRASTRUCTURE ONLY. The reference is pure Python and cannot travel to the GPU box, so its
outputs are committed here (tests/golden/*.npz, *.json) together with this script.

Run (in the build container, where /root/reference exists):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=oracle/ref_shim:/root/reference \
        python tests/golden/generate_goldens.py [board mcts_small mcts_kat mcts_noise net games sink mcts_sweep]

`oracle/ref_shim` supplies stand-ins for the three third-party imports that are not installed here
(anytree / matplotlib.pyplot / visdom). Nothing of the reference is patched
except where stated below (np.random.choice is replaced by a draw-for-draw equivalent that records the
uniform it consumed, so that sampled moves can be replayed).
"""
import json
import os
import random
import sys
from copy import copy
from functools import partial

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

from oinkoink.board import Board, make_random_ips  # noqa: E402
from oinkoink import evaluators as evl  # noqa: E402
from oinkoink.mcts import MCTS, MCTSConfig, search  # noqa: E402
from oinkoink.utils import Result  # noqa: E402

RES_CODE = {None: -1, Result.x_win: 0, Result.draw: 1, Result.o_win: 2}  # value = code * 0.5


def u64(x):
    return np.uint64(int(x) & 0xFFFFFFFFFFFFFFFF)


def mask_of(moves):
    m = 0
    for c in moves:
        m |= 1 << int(c)
    return m


def random_position(rng, plies):
    """uniformly random legal moves from the empty board; None if terminal."""
    b = Board()
    for _ in range(plies):
        b.make_move(rng.choice(sorted(b.valid_moves)))
        if b.result is not None:
            return None
    return b


def sweep_positions(seed, n, max_plies=34):
    rng = random.Random(seed)
    out = []
    while len(out) < n:
        plies = rng.randint(0, max_plies)
        b = random_position(rng, plies)
        if b is not None:
            out.append(b)
    return out


# --------------------------------------------------------------------------- board
def gen_board():
    import tests.board_test as bt
    import tests.player_test as pt

    kat = {"result_cases": [], "valid_move_cases": [], "player_cases": []}
    for o, x, a in zip(bt.pieces_1, bt.pieces_2, bt.ans):
        b = Board.from_pieces(o_pieces=o, x_pieces=x)
        assert b.result == (Result(a) if a is not None else None)
        kat["result_cases"].append({
            "o": o.astype(int).tolist(), "x": x.astype(int).tolist(), "ans": a,
            "c0": int(b.color[0]), "c1": int(b.color[1]), "age": int(b.age),
            "height": [int(h) for h in b.height]})

    # the positions of test_valid_moves live inside the function body: capture them by wrapping from_pieces
    captured = []
    orig = Board.from_pieces.__func__

    def spy(cls, o_pieces, x_pieces):
        b = orig(cls, o_pieces, x_pieces)
        captured.append((np.array(o_pieces), np.array(x_pieces), b))
        return b
    Board.from_pieces = classmethod(spy)
    try:
        bt.test_valid_moves()
    finally:
        Board.from_pieces = classmethod(orig)
    for o, x, b in captured:
        kat["valid_move_cases"].append({
            "o": o.astype(int).tolist(), "x": x.astype(int).tolist(),
            "valid": sorted(int(m) for m in b.valid_moves),
            "c0": int(b.color[0]), "c1": int(b.color[1]), "age": int(b.age)})

    for o, x, p, a in zip(pt.o_pieces, pt.x_pieces, pt.plies, pt.ans):
        b = Board.from_pieces(o_pieces=o, x_pieces=x)
        kat["player_cases"].append({
            "o": o.astype(int).tolist(), "x": x.astype(int).tolist(), "plies": p, "ans": a,
            "c0": int(b.color[0]), "c1": int(b.color[1]), "age": int(b.age)})

    kat["make_random_ips_counts"] = [len(make_random_ips(p)) for p in range(4)]
    kat["make_random_ips_2"] = sorted([int(b.color[0]), int(b.color[1])] for b in make_random_ips(2))
    with open(os.path.join(HERE, "board_kat.json"), "w") as f:
        json.dump(kat, f)

    # random playouts: every ply of 300 random games
    rng = random.Random(1)
    rows = []
    planes = []
    for g in range(300):
        b = Board()
        while True:
            fl = b.create_fliplr()
            rows.append((int(b.color[0]), int(b.color[1]), int(b.age), RES_CODE[b.result],
                         mask_of(b.valid_moves), int(fl.color[0]), int(fl.color[1]),
                         int(bool(b.symmetrical)), g))
            planes.append(np.packbits(b.to_array().astype(np.uint8).reshape(-1)))
            assert [int(h) for h in fl.height] == [7 * i + bin((int(fl.color[0]) | int(fl.color[1])) >> (7 * i) & 127).count("1") for i in range(7)]
            if b.result is not None:
                break
            mv = rng.choice(sorted(b.valid_moves))
            rows.append  # noqa
            b.make_move(mv)
    rows = np.array(rows, dtype=np.int64)
    # the move played from row i to row i+1 (same game) is recoverable from the colour difference
    np.savez_compressed(os.path.join(HERE, "board_playouts.npz"),
                        c0=rows[:, 0].astype(np.uint64), c1=rows[:, 1].astype(np.uint64),
                        age=rows[:, 2].astype(np.int8), result=rows[:, 3].astype(np.int8),
                        valid=rows[:, 4].astype(np.uint8), f0=rows[:, 5].astype(np.uint64),
                        f1=rows[:, 6].astype(np.uint64), sym=rows[:, 7].astype(np.uint8),
                        game=rows[:, 8].astype(np.int32), planes=np.stack(planes))
    print("board: %d playout rows" % len(rows))


# --------------------------------------------------------------------------- mcts
def count_nodes(node):
    n = 1
    d = 0
    for c in node.children:
        cn, cd = count_nodes(c)
        n += cn
        d = max(d, cd + 1)
    return n, d


def tree_record(tree, board_before):
    root = tree.root
    visits = np.zeros(7, np.int32)
    vsum = np.zeros(7, np.float64)
    cres = np.full(7, -2, np.int8)  # -2: no such child, -1: non-terminal, 0/1/2: result code
    for c in root.children:
        cres[c.name] = RES_CODE[c.data.board.result]
        if c.data.search_value is not None:
            visits[c.name] = c.data.search_value.visit_count
            vsum[c.name] = c.data.search_value.value_sum
    best = tree.best_move()
    nn, depth = count_nodes(root)
    return dict(
        c0=u64(board_before.color[0]), c1=u64(board_before.color[1]), age=np.int8(board_before.age),
        visits=visits, vsum=vsum, cres=cres,
        root_visits=np.int32(root.data.search_value.visit_count),
        root_vsum=np.float64(root.data.search_value.value_sum),
        root_prior=np.asarray(root.data.position_value.prior, dtype=np.float64),
        best=np.int8(best.name), best_value=np.float64(best.data.absolute_value),
        vpolicy=np.asarray(tree.get_values_policy(), np.float64),
        cpolicy=np.asarray(tree.get_visit_count_policy(), np.float64),
        nodes=np.int32(nn), depth=np.int16(depth))


def stack(recs):
    return {k: np.stack([r[k] for r in recs]) for k in recs[0]}


def _search_one(args):
    c0, c1, age, sims, pb_c_base, pb_c_init = args
    b = Board()
    b.color[0], b.color[1], b.age = c0, c1, age
    for i in range(7):
        b.height[i] = 7 * i + bin(((int(c0) | int(c1)) >> (7 * i)) & 127).count("1")
    cfg = MCTSConfig(simulations=sims, pb_c_base=pb_c_base, pb_c_init=pb_c_init)
    tree = search(cfg, b, evl.Evaluator(evl.evaluate_centre_with_prior))
    rec = tree_record(tree, b)
    rec["sims"] = np.int32(sims)
    rec["pb_c_base"] = np.float64(pb_c_base)
    rec["pb_c_init"] = np.float64(pb_c_init)
    return rec


def run_searches(jobs, procs=8):
    from multiprocessing import Pool
    with Pool(procs) as pool:
        return pool.map(_search_one, jobs, chunksize=4)


def gen_mcts_small():
    """Varied simulation counts on 240 random positions (fast)."""
    boards = sweep_positions(7, 240)
    jobs = []
    sims_list = [1, 2, 3, 7, 8, 9, 50, 200]
    for i, b in enumerate(boards):
        jobs.append((int(b.color[0]), int(b.color[1]), int(b.age), sims_list[i % len(sims_list)], 19652, 1.25))
    recs = run_searches(jobs)
    np.savez_compressed(os.path.join(HERE, "mcts_small.npz"), **stack(recs))
    print("mcts_small:", len(recs))


def gen_mcts_kat():
    """The 7 tactical positions of tests/player_test.py with the reference's exact test configuration."""
    import tests.player_test as pt
    jobs = []
    for b, p in zip(pt.boards, pt.plies):
        sims = 7 ** p + 1 if p <= 6 else 2 ** p
        jobs.append((int(b.color[0]), int(b.color[1]), int(b.age), sims, 19652, 9999))
    recs = run_searches(jobs)
    for r, a in zip(recs, pt.ans):
        assert int(r["best"]) in a
    d = stack(recs)
    d["ans_mask"] = np.array([mask_of(a) for a in pt.ans], np.uint8)
    np.savez_compressed(os.path.join(HERE, "mcts_kat.npz"), **d)
    print("mcts_kat:", len(recs))


def gen_mcts_sweep(n=10000):
    """BASELINE.json configs[1]: 10k random positions, 800 simulations, deterministic evaluator."""
    boards = sweep_positions(20261018, n)
    jobs = [(int(b.color[0]), int(b.color[1]), int(b.age), 800, 19652, 1.25) for b in boards]
    recs = run_searches(jobs)
    d = stack(recs)
    # keep the fixture small: drop what is derivable (policies are re-derived from visits/vsum/cres by the tests'
    # oracle, but keep vpolicy because its fp64 bits are part of the contract)
    for k in ("cpolicy", "pb_c_base", "pb_c_init", "sims", "root_prior"):
        d.pop(k)
    np.savez_compressed(os.path.join(HERE, "mcts_sweep_800.npz"), **d)
    print("mcts_sweep:", len(recs))


# --------------------------------------------------------------------------- noise + sampling
class Recorder:
    """Replaces np.random.gamma / np.random.choice by recording equivalents.

    choice(range(k), p=p) in legacy numpy draws ONE uniform u = random_sample() and returns
    searchsorted(cumsum(p)/cumsum(p)[-1], u, side='right'); verified below against the real function
    on the same generator state."""

    def __init__(self):
        self.noise = []
        self.uniform = []
        self._gamma = np.random.gamma
        self._choice = np.random.choice

    def gamma(self, shape, scale, size):
        g = self._gamma(shape, scale, size)
        self.noise.append(np.array(g, np.float64))
        return g

    def choice(self, a, p=None):
        st = np.random.get_state()
        ref = self._choice(a, p=p)
        np.random.set_state(st)
        u = np.random.random_sample()
        cdf = np.cumsum(np.asarray(p, np.float64))
        cdf /= cdf[-1]
        idx = int(np.searchsorted(cdf, u, side="right"))
        assert list(a)[idx] == ref
        self.uniform.append(u)
        return list(a)[idx]

    def __enter__(self):
        np.random.gamma = self.gamma
        np.random.choice = self.choice
        return self

    def __exit__(self, *a):
        np.random.gamma = self._gamma
        np.random.choice = self._choice


def play_training_game(player, rec=None):
    from oinkoink.neural.training_game import training_game
    gd = training_game(player)
    return gd


def gen_games():
    """Whole self-play games through the reference's training_game (neural/training_game.py:8-19)."""
    from oinkoink.neural.training_game import training_game
    out = {}
    # (a) deterministic: no noise, no sampling
    for gi, sims in enumerate([30, 100, 800]):
        player = MCTS("g", MCTSConfig(simulations=sims), evl.Evaluator(evl.evaluate_centre_with_prior))
        gd = training_game(player)
        out["det%d_sims" % gi] = np.int32(sims)
        out["det%d_moves" % gi] = np.array(gd.moves, np.int8)
        out["det%d_values" % gi] = np.array(gd.values, np.float64)
        out["det%d_priors" % gi] = np.array(gd.priors, np.float64)
        out["det%d_c0" % gi] = np.array([u64(b.color[0]) for b in gd.boards])
        out["det%d_c1" % gi] = np.array([u64(b.color[1]) for b in gd.boards])
        out["det%d_result" % gi] = np.float64(gd.result.value)
    # (b) AlphaZero settings with recorded randomness (MCTSConfig(sims,19652,1.25,0.3,0.25,6), training.py:209-216)
    for gi, (sims, seed) in enumerate([(60, 0), (200, 1), (800, 2), (100, 3)]):
        np.random.seed(seed)
        with Recorder() as r:
            player = MCTS("g", MCTSConfig(sims, 19652, 1.25, 0.3, 0.25, 6),
                          evl.Evaluator(evl.evaluate_centre_with_prior))
            gd = training_game(player)
        n = len(gd.moves)
        assert len(r.noise) == n
        uni = np.zeros(n, np.float64)
        uni[:len(r.uniform)] = r.uniform
        out["az%d_sims" % gi] = np.int32(sims)
        out["az%d_moves" % gi] = np.array(gd.moves, np.int8)
        out["az%d_values" % gi] = np.array(gd.values, np.float64)
        out["az%d_priors" % gi] = np.array(gd.priors, np.float64)
        out["az%d_noise" % gi] = np.stack(r.noise)
        out["az%d_uniform" % gi] = uni
        out["az%d_n_uniform" % gi] = np.int32(len(r.uniform))
        out["az%d_result" % gi] = np.float64(gd.result.value)
    np.savez_compressed(os.path.join(HERE, "games.npz"), **out)
    print("games: done")


def gen_mcts_noise():
    """Single searches with root noise on random positions; noise vectors recorded."""
    boards = sweep_positions(99, 48, max_plies=24)
    recs = []
    np.random.seed(1234)
    for i, b in enumerate(boards):
        sims = [25, 100, 400][i % 3]
        with Recorder() as r:
            cfg = MCTSConfig(sims, 19652, 1.25, 0.3, 0.25, 0)
            tree = search(cfg, b, evl.Evaluator(evl.evaluate_centre_with_prior))
        rec = tree_record(tree, b)
        rec["sims"] = np.int32(sims)
        rec["noise"] = r.noise[0]
        recs.append(rec)
    np.savez_compressed(os.path.join(HERE, "mcts_noise.npz"), **stack(recs))
    print("mcts_noise:", len(recs))


# --------------------------------------------------------------------------- net
def gen_net():
    import torch
    from oinkoink.neural.config import ModelConfig, NetConfig
    from oinkoink.neural.pytorch.model import ModelWrapper, Net
    torch.set_num_threads(4)
    ck = "/root/reference/oinkoink/data/example_net.pth"
    mw = ModelWrapper(ModelConfig(use_gpu=False), ck)
    sd = {k: v.detach().cpu().numpy() for k, v in mw.net.state_dict().items()}
    np.savez_compressed(os.path.join(HERE, "example_net_state.npz"), **sd)

    boards = [Board()] + sweep_positions(8, 1535, max_plies=40)
    with torch.no_grad():
        values, priors = mw(list(boards))
    v1, p1 = mw(boards[0])
    assert abs(float(v1[0]) - float(values[0])) < 1e-6
    out = dict(c0=np.array([u64(b.color[0]) for b in boards]), c1=np.array([u64(b.color[1]) for b in boards]),
               age=np.array([b.age for b in boards], np.int8),
               value=values.astype(np.float32), prior=priors.astype(np.float32))

    # random-init example_config net (64 filters / 6 residual / 6 fc), torch.manual_seed(0): only outputs are stored;
    # the oracle's torch restatement must reproduce the same init draw-for-draw.
    torch.manual_seed(0)
    big = Net(NetConfig(filters=64, n_fc_layers=6, n_residuals=6))
    big.eval()
    x = torch.FloatTensor(np.stack([b.to_array() for b in boards[:256]]))
    with torch.no_grad():
        bv, bp = big(x)
    out["big_value"] = bv.numpy().astype(np.float32)
    out["big_prior"] = bp.numpy().astype(np.float32)
    out["big_param_checksum"] = np.float64(sum(float(p.double().sum()) for p in big.parameters()))
    torch.manual_seed(0)
    small = Net(NetConfig())
    small.eval()
    with torch.no_grad():
        sv, sp = small(x)
    out["rand_value"] = sv.numpy().astype(np.float32)
    out["rand_prior"] = sp.numpy().astype(np.float32)
    np.savez_compressed(os.path.join(HERE, "net_outputs.npz"), **out)
    print("net: %d positions" % len(boards))


# --------------------------------------------------------------------------- generation sink
def gen_sink():
    """native_to_pytorch(..., add_fliplr=True) on a small game (neural/pytorch/data.py:78-105)."""
    from oinkoink.neural.training_game import training_game
    from oinkoink.neural.pytorch.data import native_to_pytorch
    player = MCTS("g", MCTSConfig(simulations=40), evl.Evaluator(evl.evaluate_centre_with_prior))
    games = [training_game(player) for _ in range(1)]
    data = np.sum([g.data for g in games])
    c0 = np.array([u64(b.color[0]) for b in data.boards])
    c1 = np.array([u64(b.color[1]) for b in data.boards])
    pri = np.array(data.priors, np.float64)
    bt, vt, pt_ = native_to_pytorch(list(data.boards), list(data.values), list(data.priors), add_fliplr=True)
    np.savez_compressed(os.path.join(HERE, "sink.npz"), c0=c0, c1=c1, priors=pri,
                        values=np.array(data.values, np.float64),
                        boards_t=bt.numpy().astype(np.uint8), values_t=vt.numpy(), priors_t=pt_.numpy())
    print("sink:", tuple(bt.shape))


def gen_storage():
    """The reference's own generation sink on two deterministic games: games.pkl (neural/storage.py:12-17) and the
    flip-augmented data.pth (neural/pytorch/data.py:52-64), byte-for-byte as the reference writes them."""
    import shutil
    import tempfile
    from oinkoink.neural.training_game import training_game
    from oinkoink.neural.pytorch.data import TrainingDataStorage
    games = []
    for sims in (30, 100):
        player = MCTS("g", MCTSConfig(simulations=sims), evl.Evaluator(evl.evaluate_centre_with_prior))
        games.append(training_game(player))
    d = tempfile.mkdtemp()
    TrainingDataStorage().save(games, d)
    shutil.copy(os.path.join(d, "games.pkl"), os.path.join(HERE, "games_ref.pkl"))
    shutil.copy(os.path.join(d, "data.pth"), os.path.join(HERE, "data_ref.pth"))
    shutil.rmtree(d)
    print("storage:", [len(g.moves) for g in games])


def gen_eval():
    """ModelWrapper.evaluate / evaluate_value_only (neural/pytorch/model.py:180-198,307-342) with ValueStats / PriorStats
    (neural/stats.py) on a synthetic 8-ply-shaped labelled set, example_net.pth, CPU fp32."""
    import torch
    from oinkoink.neural.config import ModelConfig
    from oinkoink.neural.pytorch.data import Connect4Dataset
    from oinkoink.neural.pytorch.model import ModelWrapper
    torch.set_num_threads(4)
    torch.manual_seed(0)
    mw = ModelWrapper(ModelConfig(use_gpu=False), "/root/reference/oinkoink/data/example_net.pth")
    rng = random.Random(8)
    boards = []
    while len(boards) < 6000:
        b = random_position(rng, 8)
        if b is not None:
            boards.append(b)
    nrng = np.random.default_rng(8)
    values = nrng.choice(np.array([0.0, 0.5, 1.0], np.float32), size=len(boards))
    priors = np.zeros((len(boards), 7), np.float32)
    for i in range(len(boards)):                     # one or two equally good moves (PriorStats accepts either)
        k = nrng.choice(7, size=1 + int(nrng.integers(0, 2)), replace=False)
        priors[i, k] = 1.0 / len(k)
    bt = torch.FloatTensor(np.stack([b.to_array() for b in boards]))
    vt, pt_ = torch.FloatTensor(values), torch.FloatTensor(priors)
    with torch.no_grad():
        vo, po = mw.net(bt)
    combined = mw.evaluate(Connect4Dataset(bt, vt, pt_), batch_size=4096, shuffle=False)
    value_only = mw.evaluate_value_only(Connect4Dataset(bt, vt, None))

    def plain(d):
        return {str(k): ({str(a): list(map(int, b)) for a, b in v.items()} if isinstance(v, dict) else float(v))
                for k, v in d.items()}
    np.savez_compressed(os.path.join(HERE, "eval_stats.npz"),
                        c0=np.array([u64(b.color[0]) for b in boards]), c1=np.array([u64(b.color[1]) for b in boards]),
                        values=values, priors=priors, value_out=vo.numpy(), prior_out=po.numpy(),
                        combined=json.dumps(plain(combined.to_dict())), combined_repr=repr(combined),
                        value_only=json.dumps(plain(value_only.to_dict())), value_only_repr=repr(value_only))
    print("eval:", repr(combined))


def gen_match():
    """Match (match.py:14-76) between two different deterministic MCTS players over every 1-ply / 2-ply opening with
    sides switched -- the shape of TrainingLoop._match (neural/training.py:176-207).  Per game: opening, move history,
    result; plus the W/D/L dictionary the reference returns."""
    from oinkoink.match import Match
    out = {}
    for mi, (plies, cfg1, cfg2) in enumerate([
            (1, MCTSConfig(simulations=200), MCTSConfig(simulations=100, pb_c_init=2.5)),
            (2, MCTSConfig(simulations=60), MCTSConfig(simulations=25, pb_c_init=0.8))]):
        p1 = MCTS("one", cfg1, evl.Evaluator(evl.evaluate_centre_with_prior))
        p2 = MCTS("two", cfg2, evl.Evaluator(evl.evaluate_centre_with_prior))
        match = Match(False, p1, p2, plies=plies, switch=True)
        starts = [(u64(g._board.color[0]), u64(g._board.color[1])) for g in match.games]
        res = match.play()
        out["m%d_plies" % mi] = np.int32(plies)
        out["m%d_cfg" % mi] = np.array([[c.simulations, c.pb_c_base, c.pb_c_init] for c in (cfg1, cfg2)], np.float64)
        out["m%d_c0" % mi] = np.array([s[0] for s in starts])
        out["m%d_c1" % mi] = np.array([s[1] for s in starts])
        out["m%d_n" % mi] = np.int32(match.n)
        hist = np.full((len(match.games), 42), -1, np.int8)
        for i, g in enumerate(match.games):
            hist[i, :len(g.move_history)] = g.move_history
        out["m%d_moves" % mi] = hist
        out["m%d_result" % mi] = np.array([g._board.result.value for g in match.games], np.float64)
        out["m%d_summary" % mi] = np.array([res["wins"], res["draws"], res["losses"], res["return"]], np.float64)
        print("match", mi, res)
    np.savez_compressed(os.path.join(HERE, "match.npz"), **out)


PARTS = dict(match=gen_match, eval=gen_eval, storage=gen_storage, board=gen_board, mcts_small=gen_mcts_small, mcts_kat=gen_mcts_kat, mcts_noise=gen_mcts_noise,
             net=gen_net, games=gen_games, sink=gen_sink, mcts_sweep=gen_mcts_sweep)

if __name__ == "__main__":
    parts = sys.argv[1:] or ["board", "mcts_small", "mcts_kat", "mcts_noise", "net", "games", "sink"]
    for p in parts:
        PARTS[p]()
