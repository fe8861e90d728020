"""`Tree`, `NodeData` and anytree-style `Node` views over a search tree that lives in the GPU node pool
(reference surface: oinkoink/tree.py:18-147).

The device stores 32-byte node records in 8-slot child blocks (csrc/c4_common.cuh); `Tree` wraps the exported pool of
one game and materialises nodes lazily.  Children are allocated eagerly on the device when a node is evaluated, but a
node only *shows* children once it has been visited twice -- exactly when the reference's lazy `expand_node` would
have created them -- so traversals see the reference's tree.
"""
import numpy as np
from scipy.special import softmax

from .board import Board
from .utils import Connect4Stats as info
from .utils import RESULT_FROM_CODE, Side, value_to_side

META_EXISTS, META_TERMINAL = 1, 2


class PositionEvaluation():
    """oinkoink/mcts.py:29-44"""

    def __init__(self, value, prior):
        self.value = value
        self.prior = prior

    def __float__(self):
        return float(self.value)

    def __str__(self):
        return str("{:.4f}".format(self.__float__()))

    def __repr__(self):
        return "position_value: " + str(self.value) + ", prior: " + str(self.prior)


class SearchEvaluation():
    """oinkoink/mcts.py:46-66"""

    def __init__(self, value_sum=0.0, visit_count=0):
        self.value_sum = value_sum
        self.visit_count = visit_count

    def add(self, value):
        self.value_sum += value
        self.visit_count += 1

    def __float__(self):
        assert self.visit_count != 0
        return float(self.value_sum / self.visit_count)

    def __str__(self):
        return str("{:.4f}".format(self.__float__()))

    def __repr__(self):
        return "value: " + str(self.__float__()) + ",  value_sum: " + str(self.value_sum) + \
            ",  visit_count: " + str(self.visit_count)


class NodeData():
    """oinkoink/tree.py:18-58"""

    def __init__(self, board, position_value=None, search_value=None):
        self.board = board
        self.valid_moves = board.valid_moves
        self.position_value = position_value
        self.search_value = search_value

    @property
    def absolute_value(self):
        if self.board.result is not None:
            return self.board.result.value
        elif self.search_value is not None:
            return float(self.search_value)
        elif self.position_value is not None:
            return float(self.position_value)
        return None

    def value(self, side):
        a = self.absolute_value
        if a is not None:
            return value_to_side(a, side)
        return 0.0  # position is unknown - assume lost

    def __str__(self):
        return "board_result: " + str(self.board.result) + ",  position_value: (" + str(self.position_value) + ")" + \
            ",  search_value: (" + str(self.search_value) + ")"

    __repr__ = __str__


class Node():
    """anytree.Node-compatible view of one pool slot: name, parent, children, is_root, data."""

    def __init__(self, tree, slot, name, parent, board):
        self._tree = tree
        self._slot = slot
        self.name = name
        self.parent = parent
        self._board = board
        self._children = None
        self._data = None

    @property
    def is_root(self):
        return self.parent is None

    def __gt__(self, other):   # oinkoink/tree.py:11-15
        return self.name > other.name

    @property
    def data(self):
        if self._data is None:
            rec = self._tree._pool[self._slot]
            sv = SearchEvaluation(float(rec["vsum"]), int(rec["visits"])) if rec["visits"] > 0 else None
            pv = None
            blk = int(rec["child_block"])
            if blk:
                kids = self._tree._pool[blk * 8: blk * 8 + 7]
                pv = PositionEvaluation(float(self._tree._pool[blk * 8 + 7]["vsum"]), kids["prior"].astype(np.float64))
            self._data = NodeData(self._board, pv, sv)
        return self._data

    @property
    def children(self):
        if self._children is None:
            rec = self._tree._pool[self._slot]
            kids = []
            blk = int(rec["child_block"])
            if blk and rec["visits"] >= 2 and not (rec["meta"] & META_TERMINAL):
                for c in range(7):
                    k = self._tree._pool[blk * 8 + c]
                    if k["meta"] & META_EXISTS:
                        b = self._board.__copy__()
                        b.make_move(c)
                        kids.append(Node(self._tree, blk * 8 + c, c, self, b))
            self._children = kids
        return tuple(self._children)

    @children.setter
    def children(self, value):
        self._children = list(value)


class Tree():
    def __init__(self, board, pool=None):
        self.side = board.player_to_move
        self._pool = pool
        self.root = Node(self, 0, 'root', None, board.__copy__())

    def get_node_value(self, node):
        return node.data.value(self.side)

    def best_move(self):
        _, child = max(((self.get_node_value(child), child) for child in self.root.children))
        return child

    def sample_value_fn(self, fn):
        values = [fn(self.get_node_value(c)) for c in self.root.children]
        probabilities = values / np.sum(values)
        idx = np.random.choice(range(len(values)), p=probabilities)
        return self.root.children[idx]

    def most_visited(self):
        _, child = max(((child.data.search_value.visit_count if child.data.search_value is not None else 0, child)
                        for child in self.root.children))
        return child

    def softmax_visit_count(self):
        visit_counts = [c.data.search_value.visit_count if c.data.search_value else 0 for c in self.root.children]
        idx = np.random.choice(range(len(visit_counts)), p=softmax(visit_counts))
        return self.root.children[idx]

    def get_values_policy(self):
        policy = np.zeros((info.width,))
        for c in self.root.children:
            policy[c.name] = self.get_node_value(c)
        self._normalise_policy(policy)
        return policy

    def get_visit_count_policy(self):
        policy = np.zeros((info.width,))
        for c in self.root.children:
            if c.data.search_value is not None:
                policy[c.name] = c.data.search_value.visit_count
        self._normalise_policy(policy)
        return policy

    def _normalise_policy(self, policy):
        s = np.sum(policy)
        if s == 0.0:
            for c in self.root.children:
                policy[c.name] = 1.0
            policy /= len(self.root.children)
        else:
            policy /= s

    def count_nodes(self):
        def rec(n):
            return 1 + sum(rec(c) for c in n.children)
        return rec(self.root)
