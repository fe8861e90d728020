"""Player protocol (oinkoink/player.py:7-35): make_move(board) -> (move, value, tree); mutates `board`."""
from .utils import Side


class BasePlayer():
    def __init__(self, name):
        self.name = name

    def __str__(self):
        return "Player: " + self.name

    def make_move(self, board):
        raise NotImplementedError


class HumanPlayer(BasePlayer):
    def make_move(self, board):
        move = -1
        while move not in board.valid_moves:
            try:
                move = int(input("Enter " + self.name + " (" + Side.as_str(board.player_to_move) + "'s) move:"))
            except ValueError:
                print("Not a valid move. Try again:")
        board.make_move(int(move))
        return move, None, None

    def __str__(self):
        return super().__str__() + ", type: Human"
