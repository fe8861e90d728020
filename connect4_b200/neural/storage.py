"""`GameStorage` (oinkoink/neural/storage.py:11-36): games.pkl + a printable last game.

The pickle is written in the REFERENCE's class namespace (`oinkoink.board.Board`,
`oinkoink.neural.training_game.GameData`, `oinkoink.utils.Result`), so the reference's own tools
(`scripts/view_games.py`, `scripts/game_stats.py`, `TrainingLoop._train`) open a generation produced here unchanged, and
`load_games` reads a games.pkl written by either side into this package's classes.  The instance state of the three
classes is identical on both sides (board.py:36-41, training_game.py:42-49, utils.py:19-22), only the module paths differ.
"""
import importlib
import pickle
import sys
import threading
import types
from enum import Enum

from ..board import Board
from ..utils import Result, Side
from . import training_game as _tg

# (reference module, name) <-> class of this package
_CLASS_MAP = {
    ("oinkoink.board", "Board"): Board,
    ("oinkoink.neural.training_game", "GameData"): _tg.GameData,
    ("oinkoink.neural.training_game", "TrainingData"): _tg.TrainingData,
    ("oinkoink.utils", "Result"): Result,
    ("oinkoink.utils", "Side"): Side,
}
_LOCK = threading.Lock()
_MISSING = object()


class _ReferenceUnpickler(pickle.Unpickler):
    def find_class(self, module, name):
        cls = _CLASS_MAP.get((module, name))
        if cls is not None:
            return cls
        return super().find_class(module, name)


def load_games(file_name):
    """games.pkl (written by the reference's GameStorage.save or by this package) -> List[GameData] of this package."""
    with open(file_name, 'rb') as f:
        return _ReferenceUnpickler(f).load()


class _ReferenceNamespace():
    """While active, `oinkoink.*` resolves for the pickler: the real reference if it is importable, otherwise
    stand-in modules holding state-compatible stub classes (removed again on exit)."""

    def __enter__(self):
        self.fake = []
        self.cls = {}
        self.saved = []                      # (module object, attribute, previous value or _MISSING)
        before = set(sys.modules)
        try:
            for (mod, name) in _CLASS_MAP:
                self.cls[(mod, name)] = getattr(importlib.import_module(mod), name)
            return self
        except Exception:
            # a partly importable reference (e.g. its `anytree` dependency missing): drop what this attempt imported, so
            # that no real module is left half-initialised, and fall back to stand-ins
            for m in set(sys.modules) - before:
                if m == "oinkoink" or m.startswith("oinkoink."):
                    sys.modules.pop(m, None)
            self.cls = {}
        for (mod, name), ours in _CLASS_MAP.items():
            parts = mod.split(".")
            for i in range(1, len(parts) + 1):
                m = ".".join(parts[:i])
                if m not in sys.modules:
                    sys.modules[m] = types.ModuleType(m)
                    self.fake.append(m)
                    if i > 1:
                        setattr(sys.modules[".".join(parts[:i - 1])], parts[i - 1], sys.modules[m])
            if issubclass(ours, Enum):
                stub = ours.__bases__[0](name, {k: v.value for k, v in ours.__members__.items()}, module=mod)
            else:
                stub = type(name, (object,), {"__module__": mod, "__qualname__": name})
            target = sys.modules[mod]
            self.saved.append((target, name, getattr(target, name, _MISSING)))     # a real module loaded earlier keeps its
            setattr(target, name, stub)                                             # own class after __exit__
            self.cls[(mod, name)] = stub
        return self

    def __exit__(self, *a):
        for target, name, prev in reversed(self.saved):
            if prev is _MISSING:
                if hasattr(target, name):
                    delattr(target, name)
            else:
                setattr(target, name, prev)
        for m in reversed(self.fake):
            sys.modules.pop(m, None)

    # ---- object graph conversion (GameData / TrainingData / Board / Result -> reference-namespace twins)
    def convert(self, obj):
        for key, ours in _CLASS_MAP.items():
            if type(obj) is ours:
                theirs = self.cls[key]
                if isinstance(obj, Enum):
                    return theirs(obj.value)
                twin = theirs.__new__(theirs)
                twin.__dict__.update({k: self.convert(v) for k, v in obj.__dict__.items()})
                return twin
        if type(obj) is list:
            return [self.convert(v) for v in obj]
        if type(obj) is tuple:
            return tuple(self.convert(v) for v in obj)
        if type(obj) is dict:
            return {k: self.convert(v) for k, v in obj.items()}
        return obj


def dump_games(games, file_name, protocol=pickle.DEFAULT_PROTOCOL):
    """Write `games` (List[GameData]) as the reference's games.pkl (storage.py:12-17)."""
    with _LOCK, _ReferenceNamespace() as ns:
        twins = ns.convert(games)
        with open(file_name, 'wb') as f:
            pickle.dump(twins, f, protocol)


class GameStorage():
    def save(self, games, folder_path):
        dump_games(games, folder_path + '/games.pkl')
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
