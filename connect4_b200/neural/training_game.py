"""Self-play game records with the reference's shapes (oinkoink/neural/training_game.py:8-72): `GameData` is what a
generation is made of (and what `games.pkl` stores), `TrainingData` the flat (boards, values, priors) view the sink
consumes.  On the device a generation is an array of 64-byte position records; `games_from_records` turns it into
`GameData` objects."""
from copy import copy
from typing import List, Sequence

import numpy as np

from ..board import Board
from ..utils import RESULT_FROM_CODE


class GameData():
    """instance state (and its order, which the pickle stream follows): result, moves, boards, values, priors"""

    def __init__(self):
        self.result = None          # utils.Result once the game is over
        self.moves = []             # column played at every ply
        self.boards = []            # position BEFORE each move
        self.values = []            # the chosen child's absolute value after the search (None if it has none)
        self.priors = []            # Tree.get_values_policy() of each search: the policy training target

    def add_move(self, board, move, value, prior):
        for column, item in ((self.moves, move), (self.boards, board), (self.values, value), (self.priors, prior)):
            column.append(item)

    def create_training_values(self):
        """every position of the game is labelled with the game's result (training_game.py:57-60)"""
        return len(self.values) * [self.result.value]

    @property
    def data(self):
        assert self.result is not None
        return TrainingData(self.boards, self.create_training_values(), self.priors)

    def __str__(self):
        return "Result: {}, Moves: {}".format(self.result, self.moves)

    @classmethod
    def from_records(cls, recs):
        """recs: the contiguous, ply-ordered device records of ONE game (engine.RECORD_DTYPE)."""
        game = cls()
        for r in recs:
            value = float(r["search_value"])
            game.add_move(Board.from_bitboards(int(r["c0"]), int(r["c1"])), int(r["move"]),
                          None if np.isnan(value) else value, r["policy"].astype(np.float64))
        game.result = RESULT_FROM_CODE[int(recs[-1]["result"])]
        return game


class TrainingData:
    """three parallel lists; `+` concatenates, and 0 + data = data so that sum() over games works"""

    def __init__(self, boards: List[Board], values: List[float], priors: List[Sequence[float]]):
        self.boards, self.values, self.priors = boards, values, priors

    def __add__(self, other: 'TrainingData'):
        return TrainingData(*(mine + theirs for mine, theirs in
                              ((self.boards, other.boards), (self.values, other.values), (self.priors, other.priors))))

    def __radd__(self, other):
        return self if other == 0 else NotImplemented

    def __repr__(self):
        return str(list(zip(self.boards, self.values, self.priors)))


def training_game(player):
    """One self-play game with `player` on both sides (training_game.py:8-19); every move is a device search
    (MCTS.make_move), and the position is logged as it was BEFORE the move."""
    board, game = Board(), GameData()
    while board.result is None:
        before = copy(board)
        move, value, tree = player.make_move(board)
        game.add_move(before, move, value, tree.get_values_policy())
    game.result = board.result
    return game


def games_from_records(records):
    """Split a generation's record array into per-game GameData (ordered by global game id)."""
    if len(records) == 0:
        return []
    recs = records[np.lexsort((records["ply"], records["game_id"]))]
    cuts = np.flatnonzero(np.diff(recs["game_id"])) + 1
    return [GameData.from_records(chunk) for chunk in np.split(recs, cuts)]
