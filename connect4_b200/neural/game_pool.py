"""Self-play generation on the device game pool -- the replacement for the reference's concurrency runtime
(`game_pool`, oinkoink/neural/game_pool.py:15-49; `InferenceServer`, neural/inference_server.py:15-76;
`TrainingLoop._generate_games`, neural/training.py:99-145).

The reference keeps 10 processes x 20 threads of sequential searches busy and funnels single-board requests through
pipes to one inference process.  Here one CUDA context holds `concurrent_games` trees; each lock-step pass advances
every tree to its next leaf, evaluates all leaves in one network launch and backs the answers up.  Finished games are
re-seeded at once, so a pass always carries a full batch.
"""
from typing import List

import numpy as np

from ..engine import Engine, augment_pack
from .training_game import GameData, games_from_records


class SelfPlayPool():
    def __init__(self, evaluator, mcts_config, concurrent_games=4096, seed=0, device=None):
        """evaluator: a ModelWrapper (network) or the string 'centre' (deterministic evaluate_centre_with_prior)."""
        self.mcts_config = mcts_config
        self.engine = Engine(concurrent_games, mcts_config, device=device)
        self.kind = "centre" if isinstance(evaluator, str) else "net"
        if self.kind == "net":
            self.engine.set_net(evaluator)
        self.model = evaluator
        noisy = bool(mcts_config.root_dirichlet_alpha and mcts_config.root_exploration_fraction) or \
            mcts_config.num_sampling_moves > 0
        self.engine.set_rng("philox" if noisy else "none", seed=seed)

    def generate_records(self, n_games, game_id_base=0, game_id_stride=1, start=None, to_host=True):
        """Play n_games games; returns the 64-byte position records (numpy, engine.RECORD_DTYPE)."""
        return self.engine.selfplay(n_games, self.kind, game_id_base, game_id_stride, start, to_host)

    def generate(self, n_games, **kw) -> List[GameData]:
        """`game_pool(...) -> List[GameData]` of the reference."""
        return games_from_records(self.generate_records(n_games, **kw))

    def last_dataset(self):
        """(boards [2P,3,6,7], values [2P], priors [2P,7]) float32 CUDA tensors of the last generation with the
        reference's flip augmentation (neural/pytorch/data.py:78-105), packed on the device."""
        return augment_pack(self.engine.last_records_device)

    def stream(self, stop_games=0, max_ms=0.0, reset=False, cold_memo=False):
        """continuous self-play with immediate re-seeding (BASELINE configs[2]); see c4_selfplay_stream"""
        return self.engine.stream(self.kind, stop_games, max_ms, reset, cold_memo)

    def throughput(self, iterations):
        """steady-state measurement: see c4_selfplay_bench in include/c4b200.h"""
        return self.engine.bench(iterations, self.kind)


def game_pool(evaluator, n_threads: int, mcts_config, n_games: int, seed=0) -> List[GameData]:
    """Signature-compatible stand-in for neural/game_pool.py:15-18 (`conn_list` -> the evaluator itself;
    `n_threads` -> concurrent game slots)."""
    return SelfPlayPool(evaluator, mcts_config, concurrent_games=max(1, n_threads), seed=seed).generate(n_games)
