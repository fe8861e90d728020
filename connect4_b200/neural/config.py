"""Configuration objects with the attribute names and defaults of the reference (oinkoink/neural/config.py:7-81), so that
an `example_config.py` written for the reference constructs them unchanged.  Each class is a plain attribute bag filled
from a defaults table; unknown keywords are rejected like a wrong argument name would be."""
import os


class _Config():
    """positional arguments follow the order of DEFAULTS (the reference's constructor signatures)"""
    DEFAULTS = ()

    def __init__(self, *args, **kwargs):
        names = [n for n, _ in self.DEFAULTS]
        if len(args) > len(names):
            raise TypeError("{} takes at most {} arguments".format(type(self).__name__, len(names)))
        given = dict(zip(names, args))
        for k in kwargs:
            if k not in names:
                raise TypeError("{} got an unexpected keyword argument '{}'".format(type(self).__name__, k))
            if k in given:
                raise TypeError("{} got multiple values for argument '{}'".format(type(self).__name__, k))
        given.update(kwargs)
        for name, default in self.DEFAULTS:
            value = given.get(name, default)
            setattr(self, name, value() if callable(value) and name not in given else value)


class NetConfig(_Config):
    DEFAULTS = (("channels", 3), ("filters", 32), ("n_fc_layers", 4), ("n_residuals", 3))


class ModelConfig(_Config):
    DEFAULTS = (("net_config", NetConfig), ("weight_decay", 1e-4), ("momentum", 0.9), ("initial_lr", 0.01),
                ("milestones", lambda: [100, 300, 600]), ("gamma", 0.1), ("batch_size", 4096), ("n_training_epochs", 5),
                ("use_gpu", True))

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.net_config is None:
            self.net_config = NetConfig()
        self.milestones = list(self.milestones)


class StorageConfig(_Config):
    DEFAULTS = (("save_dir", lambda: os.path.expanduser('~')), ("data_dir", None))


class AlphaZeroConfig(_Config):
    # game_processes x game_threads was the reference's concurrency (10 x 20 = 200 games in flight); on the GPU the
    # equivalent knob is `concurrent_games`, the slots of the device game pool
    DEFAULTS = (("model_config", ModelConfig), ("storage_config", StorageConfig), ("game_processes", 10), ("game_threads", 20),
                ("simulations", 800), ("pb_c_base", 19652), ("pb_c_init", 1.25), ("root_dirichlet_alpha", 0.3),
                ("root_exploration_fraction", 0.25), ("num_sampling_moves", 6), ("n_eval", 1), ("n_training_games", 1200),
                ("use_pytorch", True), ("enable_gpu", True), ("visdom_enabled", False), ("concurrent_games", 4096))

    def __init__(self, *args, **kwargs):
        super().__init__(*args, **kwargs)
        if self.model_config is None:
            self.model_config = ModelConfig()
        if self.storage_config is None:
            self.storage_config = StorageConfig()

    def mcts_config(self, training=True):
        """the search settings of self-play (training) and of evaluation matches (TrainingLoop._create_alpha_zero_config,
        oinkoink/neural/training.py:209-223): matches run without root noise and without sampled moves"""
        from ..mcts import MCTSConfig
        noise = (self.root_dirichlet_alpha, self.root_exploration_fraction, self.num_sampling_moves) if training else (0.0, 0.0, 0)
        return MCTSConfig(self.simulations, self.pb_c_base, self.pb_c_init, *noise)
