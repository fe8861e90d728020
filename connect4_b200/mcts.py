"""`MCTSConfig`, `MCTS`, `search` with the reference's surface (oinkoink/mcts.py:13-121); the search itself runs in
the CUDA engine (csrc/c4_search.cu).  There is no host implementation of the search: without the CUDA library or a
GPU these calls raise.

Evaluator dispatch (see evaluators.Evaluator.device_kind):
  Evaluator(evaluate_centre_with_prior)            -> evaluator fused into the tree kernel, one launch per search
  Evaluator(partial(evaluate_nn, model=ModelWrapper)) -> tree kernel + CUDA network kernel, all on the device
  any other callable `evaluator(board) -> (value, prior)` -> tree kernels on the device, the callable on the host
"""
from typing import Callable, List, Tuple

import collections

import numpy as np

from .board import Board
from .evaluators import Evaluator
from .player import BasePlayer
from .tree import PositionEvaluation, SearchEvaluation, Tree  # noqa: F401  (re-exported like the reference)
from .utils import Connect4Stats as info


class MCTSConfig():
    def __init__(self, simulations: int, pb_c_base: int = 19652, pb_c_init: float = 1.25,
                 root_dirichlet_alpha: float = 0.0, root_exploration_fraction: float = 0.0, num_sampling_moves=0):
        self.simulations = simulations
        self.pb_c_base = pb_c_base
        self.pb_c_init = pb_c_init
        self.root_dirichlet_alpha = root_dirichlet_alpha
        self.root_exploration_fraction = root_exploration_fraction
        self.num_sampling_moves = num_sampling_moves


_ENGINES = collections.OrderedDict()          # (capacity, simulations, device, network tag) -> Engine, LRU first
_MAX_ENGINES = 4


def _engine(config, n, tag=None):
    """Engines are cached per (capacity, simulations, CUDA device, network): creating one allocates the node pool and the
    evaluation memo, and two networks that alternate (a Match of a new against an old network, the reference's `_match`,
    neural/training.py:176-207) each keep their memo.  The cache is a small LRU: an evicted engine is closed at once (node
    pool and memo returned to the device), so a training loop that brings a new network every generation holds at most
    _MAX_ENGINES contexts, and no cached engine keeps a reference to a network (`tag` is only a number; c4_ctx_set_net
    compares the network's unique id and empties the memo if a recycled tag ever meets another network)."""
    import torch
    from . import _lib
    from .engine import Engine
    _lib.require_gpu()                        # no CPU fallback: fail here, before anything touches the CUDA runtime
    cap = 1
    while cap < n:
        cap *= 2
    key = (cap, int(config.simulations), torch.cuda.current_device(), tag)
    eng = _ENGINES.pop(key, None)
    if eng is None:
        while len(_ENGINES) >= _MAX_ENGINES:
            _ENGINES.popitem(last=False)[1].close()
        eng = Engine(cap, config)
    _ENGINES[key] = eng                       # most recently used last
    eng.set_config(config)
    return eng


def release_engines():
    """close every cached engine (device memory of the node pools and evaluation memos)"""
    while _ENGINES:
        _ENGINES.popitem()[1].close()


def _host_evaluator(evaluator):
    def batch(c0, c1):
        values, priors = [], []
        for a, b in zip(c0, c1):
            v, p = evaluator(Board.from_bitboards(int(a), int(b)))
            values.append(float(np.asarray(v).reshape(-1)[0]))
            priors.append(np.asarray(p))
        pr = np.stack(priors)
        return np.array(values, np.float64), (pr if pr.dtype == np.float32 else pr.astype(np.float64))
    return batch


def device_kind(evaluator):
    """('centre' | 'net' | 'external', ModelWrapper or None).  A bare ModelWrapper is accepted as an evaluator like in
    the reference's `_match` (neural/training.py:189-195)."""
    if isinstance(evaluator, Evaluator):
        return evaluator.device_kind()
    if hasattr(evaluator, "c4_net"):
        return "net", evaluator
    return "external", None


def search_batch(config: MCTSConfig, boards: List[Board], evaluator, noise=None):
    """search() for many root positions at once (one GPU warp per tree). Returns (engine, trees-as-readout dict)."""
    kind, model = device_kind(evaluator)
    eng = _engine(config, len(boards), id(model) if model is not None else None)
    c0 = np.array([int(b.color[0]) for b in boards], np.uint64)
    c1 = np.array([int(b.color[1]) for b in boards], np.uint64)
    if noise is not None:
        eng.set_rng("injected", noise=np.asarray(noise, np.float64).reshape(len(boards), 1, 7),
                    uniform=np.zeros((len(boards), 1)))
    else:
        eng.set_rng("none")
    eng.begin(c0, c1)
    if kind == "centre":
        eng.run("centre")
    elif kind == "net":
        eng.set_net(model)
        try:
            eng.run("net")
        finally:
            eng.net = None                        # the cache must not keep the caller's network alive
    else:
        eng.run_external(_host_evaluator(evaluator))
    return eng


def search(config: MCTSConfig, board: Board, evaluator: Callable[[Board], Tuple[float, List[float]]]):
    """oinkoink/mcts.py:94-121. Root noise (mcts.py:171-181) is drawn here with np.random.gamma -- the same call, on
    the same global numpy stream, as the reference -- and injected into the device search."""
    if board.result is not None:
        raise ValueError("search() on a finished game (the reference fails here too: mcts.py:102)")
    noise = None
    if config.root_dirichlet_alpha and config.root_exploration_fraction:
        noise = np.random.gamma(config.root_dirichlet_alpha, 1, info.width)
    eng = search_batch(config, [board], evaluator, None if noise is None else noise[None, :])
    return Tree(board, eng.export_tree(0))


class MCTS(BasePlayer):
    def __init__(self, name: str, config: MCTSConfig, evaluator: Evaluator):
        super().__init__(name)
        self.config = config
        self.evaluator = evaluator

    def make_move(self, board):
        """oinkoink/mcts.py:78-88: search, choose (sample v^2 in the opening, else best value), play it on `board`."""
        tree = search(self.config, board, self.evaluator)
        if board.age < self.config.num_sampling_moves:
            child = tree.sample_value_fn(lambda x: x ** 2)
        else:
            child = tree.best_move()
        board.make_move(child.name)
        return child.name, child.data.absolute_value, tree

    def __str__(self):
        return super().__str__() + ", type: Computer"
