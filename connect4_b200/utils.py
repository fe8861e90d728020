"""Board dimensions, sides, results (names and values of oinkoink/utils.py:4-34) and their encodings on the C ABI."""
from enum import Enum, IntEnum


class Connect4Stats():
    width, height = 7, 6
    area = width * height


class Side(IntEnum):
    """o moves on even ages, x on odd ones (board.py:84-86)"""
    o, x = 0, 1

    @classmethod
    def as_str(cls, side):
        return cls(side).name


class Result(Enum):
    """the value of a finished game from o's point of view"""
    x_win = 0.0
    draw = 0.5
    o_win = 1.0


# result codes of the C ABI (include/c4b200.h): -1 = game running, otherwise value = code * 0.5
CODE_FROM_RESULT = {None: -1, Result.x_win: 0, Result.draw: 1, Result.o_win: 2}
RESULT_FROM_CODE = {code: result for result, code in CODE_FROM_RESULT.items()}


def same_side(result: Result, side: Side) -> bool:
    """did `side` win?"""
    return {Result.o_win: Side.o, Result.x_win: Side.x}.get(result) == side


def value_to_side(value: float, side: Side) -> float:
    """an o-perspective value seen by `side` (utils.py:33-34)"""
    return (1.0 - value) if side == Side.x else value
