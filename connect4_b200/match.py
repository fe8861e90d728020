"""`Match` (oinkoink/match.py:14-76): every n-ply opening, optionally replayed with sides switched; W/D/L + return.

The reference plays the games one after the other (or in a process pool).  Here, when both players are deterministic
`MCTS` players whose evaluators run on the device (centre evaluator or a network), ALL games advance together: per
round one batched device search for every game in which player 1 is to move and one for player 2's games
(`mcts.search_batch`, one warp per game), the chosen moves are played and the round repeats until every game has a
result.  Games never interact, so this is the same match move for move (tests/test_gpu_match.py compares every game with
the reference's).  Any other player (human, host evaluator, root noise or sampled moves) takes the sequential path.
"""
from copy import copy

import numpy as np

from .board import make_random_ips
from .game import Game
from .mcts import MCTS, device_kind, search_batch


def _batchable(player):
    if not isinstance(player, MCTS):
        return False
    c = player.config
    if c.num_sampling_moves or (c.root_dirichlet_alpha and c.root_exploration_fraction):
        return False                       # host RNG semantics (np.random) are kept by the one-game path
    return device_kind(player.evaluator)[0] in ("centre", "net")


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
        self._display = display

    def play_batched(self):
        """all games in lock step on the device; returns the per-game result values (o's perspective)"""
        games = self.games
        live = [i for i, g in enumerate(games) if g._board.result is None]
        while live:
            for role in (self._player_1, self._player_2):
                # games of this round in which a copy of `role` is to move (copies share config and evaluator)
                idx = [i for i in live if games[i]._board.result is None and
                       games[i].player_to_move().evaluator is role.evaluator and
                       games[i].player_to_move().config is role.config]
                if not idx:
                    continue
                eng = search_batch(role.config, [games[i]._board for i in idx], role.evaluator)
                best = eng.readout(len(idx))["best"]
                for i, mv in zip(idx, best.tolist()):
                    if mv < 0:      # the root was never expanded (simulations == 0); Tree.best_move fails there too
                        raise ValueError("search returned no move for a running game (simulations must be >= 1)")
                    games[i]._board.make_move(int(mv))
                    games[i].move_history = np.append(games[i].move_history, np.uint8(mv))
            live = [i for i in live if games[i]._board.result is None]
        return [g._board.result.value for g in games]

    def play(self, agents=1):
        # `agents` selected a process pool in the reference (match.py:72-76); the device needs no host parallelism
        if not self._display and _batchable(self._player_1) and _batchable(self._player_2):
            results = self.play_batched()
        else:
            results = [g.play().value for g in self.games]
        results = np.array(results, dtype='f')
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
