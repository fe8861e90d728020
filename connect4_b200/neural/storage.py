"""`GameStorage` (oinkoink/neural/storage.py:11-36): games.pkl + a printable last game."""
import pickle

from ..board import Board


class GameStorage():
    def save(self, games, folder_path):
        with open(folder_path + '/games.pkl', 'wb') as f:
            pickle.dump(games, f)
        self.last_game = games[-1]

    def last_game_str(self):
        return game_str(self.last_game.moves, self.last_game.values, self.last_game.priors)


def game_str(moves, values, policies):
    board = Board()
    out = str(board)
    for move, value, policy in zip(moves, values, policies):
        board.make_move(move)
        out += '\nMove: {}  Value: {} Policy: {}\n{}'.format(move, value, policy, board)
    return out
