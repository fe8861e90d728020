"""connect4_b200 -- the self-play generation hot path of willis-richard/connect4 (`oinkoink`) on B200 (sm_100a).

Python surface mirrors the reference for this path only: board.Board, game.Game, match.Match, player.*, mcts.MCTS /
MCTSConfig / search, tree.Tree / NodeData, evaluators.*, neural.model.ModelWrapper, neural.training_game,
neural.game_pool.  All compute goes through libc4b200.so (include/c4b200.h); there is no CPU fallback.
"""
__version__ = "0.1.0"
