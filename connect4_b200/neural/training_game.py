"""`training_game`, `GameData`, `TrainingData` (oinkoink/neural/training_game.py:8-72)."""
from copy import copy
from typing import List, Sequence

import numpy as np

from ..board import Board
from ..utils import RESULT_FROM_CODE


def training_game(player):
    """One self-play game with `player` on both sides; every move is a device search (MCTS.make_move)."""
    board = Board()
    game_data = GameData()
    while board.result is None:
        board_copy = copy(board)
        move, value, tree = player.make_move(board)
        prior = tree.get_values_policy()
        game_data.add_move(board_copy, move, value, prior)
    game_data.result = board.result
    return game_data


class TrainingData:
    def __init__(self, boards: List[Board], values: List[float], priors: List[Sequence[float]]):
        self.boards = boards
        self.values = values
        self.priors = priors

    def __add__(self, other: 'TrainingData'):
        return TrainingData(self.boards + other.boards, self.values + other.values, self.priors + other.priors)

    def __radd__(self, other):          # lets sum() / np.sum() start from 0
        return self if other == 0 else NotImplemented

    def __repr__(self):
        return str([(b, v, p) for b, v, p in zip(self.boards, self.values, self.priors)])


class GameData():
    def __init__(self):
        self.result = None
        self.moves = []
        self.boards = []
        self.values = []
        self.priors = []

    def add_move(self, board, move, value, prior):
        self.moves.append(move)
        self.boards.append(board)
        self.values.append(value)
        self.priors.append(prior)

    def create_training_values(self):
        return [self.result.value] * len(self.values)

    @property
    def data(self):
        assert self.result is not None
        return TrainingData(self.boards, self.create_training_values(), self.priors)

    def __str__(self):
        return "Result: {}, Moves: {}".format(self.result, self.moves)

    @classmethod
    def from_records(cls, recs):
        """recs: the contiguous, ply-ordered device records of ONE game (engine.RECORD_DTYPE)."""
        g = cls()
        for r in recs:
            sv = float(r["search_value"])
            g.add_move(Board.from_bitboards(int(r["c0"]), int(r["c1"])), int(r["move"]),
                       None if np.isnan(sv) else sv, r["policy"].astype(np.float64))
        g.result = RESULT_FROM_CODE[int(recs[-1]["result"])]
        return g


def games_from_records(records):
    """Split a generation's record array into per-game GameData (ordered by global game id)."""
    if len(records) == 0:
        return []
    order = np.lexsort((records["ply"], records["game_id"]))
    recs = records[order]
    cuts = np.flatnonzero(np.diff(recs["game_id"])) + 1
    return [GameData.from_records(chunk) for chunk in np.split(recs, cuts)]
