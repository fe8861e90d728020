"""`Game` (oinkoink/game.py:8-40): the two players take turns on one board until it carries a result.  Every
`make_move` of an `MCTS` player is a device search; `Match` advances many games in lock step instead (match.py)."""
import numpy as np

from .utils import Side


class Game():
    def __init__(self, display, player_o, player_x, board):
        self.display = display
        self._player_o, self._player_x = player_o, player_x
        self._board = board
        self.move_history = np.empty((0,), dtype='uint8')

    def player_to_move(self):
        return self._player_x if self._board.player_to_move == Side.x else self._player_o

    def _show(self, *lines):
        if self.display:
            for line in lines:
                print(line)

    def _report(self, player, move, value, tree):
        if tree is None:
            return "{} selected move: {}".format(player.name, move)
        return "{} selected move: {}, value: {}, prior: {}".format(player.name, move, value, tree.get_visit_count_policy())

    def play(self):
        board = self._board
        self._show("Game between {}  and  {}".format(self._player_o, self._player_x), board)
        while board.result is None:
            mover = self.player_to_move()
            move, value, tree = mover.make_move(board)
            self.move_history = np.append(self.move_history, move)
            if self.display:
                self._show(self._report(mover, move, value, tree), board)
        return board.result
