"""The player protocol of the path (oinkoink/player.py:7-35): `make_move(board) -> (move, value, tree)`, which plays the
move on `board` in place.  `MCTS` (mcts.py) is the device-backed implementation; the two classes here only fix the
protocol and give an interactive opponent for `Game`."""
from .utils import Side


class BasePlayer():
    kind = None                                # suffix of the printed form, set by subclasses

    def __init__(self, name):
        self.name = name

    def make_move(self, board):
        """must return (column, value or None, tree or None) after calling board.make_move(column)"""
        raise NotImplementedError

    def __str__(self):
        text = "Player: " + self.name
        return text if self.kind is None else "{}, type: {}".format(text, self.kind)


class HumanPlayer(BasePlayer):
    kind = "Human"

    def _prompt(self, board):
        side = Side.as_str(board.player_to_move)
        return "Enter {} ({}'s) move:".format(self.name, side)

    def make_move(self, board):
        legal = board.valid_moves
        while True:
            answer = input(self._prompt(board))
            if answer.strip().lstrip("+-").isdigit() and int(answer) in legal:
                column = int(answer)
                break
            if not answer.strip().lstrip("+-").isdigit():
                print("Not a valid move. Try again:")
        board.make_move(column)
        return column, None, None
