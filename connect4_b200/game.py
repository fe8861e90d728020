"""`Game` (oinkoink/game.py:8-40): two players alternate make_move on one board until it has a result."""
import numpy as np

from .utils import Side


class Game():
    def __init__(self, display, player_o, player_x, board):
        self.display = display
        self._player_o = player_o
        self._player_x = player_x
        self._board = board
        self.move_history = np.empty((0,), dtype='uint8')

    def player_to_move(self):
        return self._player_o if self._board.player_to_move == Side.o else self._player_x

    def play(self):
        if self.display:
            print("Game between", self._player_o, " and ", self._player_x)
            print(self._board)
        while self._board.result is None:
            player = self._player_o if self._board.player_to_move == Side.o else self._player_x
            move, value, tree = player.make_move(self._board)
            if self.display:
                if tree is None:
                    print("{} selected move: {}".format(player.name, move))
                else:
                    print("{} selected move: {}, value: {}, prior: {}".format(
                        player.name, move, value, tree.get_visit_count_policy()))
                print(self._board)
            self.move_history = np.append(self.move_history, move)
        return self._board.result
