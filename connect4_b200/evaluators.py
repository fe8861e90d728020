"""Evaluator protocol (oinkoink/evaluators.py:9-63): `evaluator(board) -> (value, prior[7])`, value absolute
(o's perspective) in [0, 1].

`evaluate_centre_with_prior` and `evaluate_nn` are MARKERS as well as callables: when an `Evaluator` wrapping one of
them is handed to `MCTS`, the search runs entirely on the device (the centre evaluator is fused into the tree kernel;
the network is the CUDA tower).  Any other callable is driven through the external-evaluator stepping interface of the
engine (c4_search_pending / c4_search_supply): tree work on the GPU, your function on the host.
Called directly they evaluate one board on the device.
"""
from copy import deepcopy
from functools import partial
from typing import Callable, Dict, Optional, Tuple

import numpy as np

from .board import Board, BoardBatch
from .utils import Connect4Stats as info


class Evaluator():
    def __init__(self, evaluate_fn: Callable, position_table: Optional[Dict[Tuple, Tuple]] = None,
                 store_position: Optional[bool] = True):
        self.evaluate_fn = evaluate_fn
        self.position_table = {} if position_table is None else position_table
        self.store_position = store_position

    def __call__(self, board: Board):
        key = (int(board.color[0]), int(board.color[1]))
        position_eval = self.position_table.get(key)
        if position_eval is None:
            position_eval = self.evaluate_fn(board)
            if self.store_position:
                self.position_table[key] = position_eval
        return deepcopy(position_eval)

    # ---- how MCTS should run this evaluator
    def device_kind(self):
        fn = self.evaluate_fn
        if fn is evaluate_centre_with_prior:
            return "centre", None
        if isinstance(fn, partial) and fn.func is evaluate_nn:
            model = fn.keywords.get("model", fn.args[0] if fn.args else None)
            if hasattr(model, "c4_net"):
                return "net", model
        return "external", None


def evaluate_centre(board: Board):
    """0.5 + (sum_o grid - sum_x grid) / 96 (evaluators.py:28-33), computed by the device bitboard engine."""
    return float(BoardBatch.from_boards([board]).evaluate_centre().cpu().numpy()[0])


def evaluate_centre_with_prior(board: Board):
    return evaluate_centre(board), np.ones((info.width,), dtype=float) / info.width


def evaluate_nn(board: Board, model):
    value, prior = model(board)
    return float(np.asarray(value).reshape(-1)[0]), prior
