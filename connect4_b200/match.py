"""`Match` (oinkoink/match.py:14-76): every n-ply opening, optionally replayed with sides switched; W/D/L + return."""
from copy import copy

import numpy as np

from .board import make_random_ips
from .game import Game


class Match():
    def __init__(self, display, player_1, player_2, plies=0, switch=False):
        self._player_1 = player_1
        self._player_2 = player_2
        ips = make_random_ips(plies)
        self.games = [Game(display, copy(player_1), copy(player_2), board) for board in ips]
        self.n = len(self.games)
        if switch:
            self.games += [Game(display, copy(player_2), copy(player_1), copy(board)) for board in ips]
        self.switch = switch

    def play(self, agents=1):
        # `agents` selected a process pool in the reference (match.py:72-76); the GPU engine needs no host
        # parallelism, so games are simply played in order.
        results = np.array([g.play().value for g in self.games], dtype='f')
        if self.switch:   # results of the games where player_2 moved first are seen from player_1's side
            results[self.n:] *= -1.0
            results[self.n:] += 1.0
        wins = np.sum(results == 1)
        draws = np.sum(results == 0.5)
        losses = np.sum(results == 0)
        return_ = (1.0 * wins + 0.5 * draws) / (wins + draws + losses)
        print("The results for {} vs {} are: {} wins, {} draws, {} losses, {:.3f} return".format(
            self._player_1.name, self._player_2.name, wins, draws, losses, return_))
        return {'wins': wins, 'draws': draws, 'losses': losses, 'return': return_}
