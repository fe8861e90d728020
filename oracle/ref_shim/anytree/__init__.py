"""Minimal stand-in for the third-party `anytree` package (not installed here, no network).

TEST INFRASTRUCTURE ONLY: lets the unmodified reference at /root/reference be imported in the
build container so that golden vectors can be generated (tests/golden/generate_goldens.py).
Never imported by the product package.

Semantics used by the reference (oinkoink/tree.py:4-15,137; tests/player_test.py:165-167):
insertion-ordered children, `.parent` settable, `.children` readable (tuple) and assignable,
`.is_root`, arbitrary keyword attributes, and `Node.__gt__` monkey-patched at import.
"""


class Node:
    def __init__(self, name, parent=None, children=None, **kwargs):
        self.name = name
        self._parent = None
        self._children = []
        for k, v in kwargs.items():
            setattr(self, k, v)
        if parent is not None:
            self.parent = parent
        if children is not None:
            self.children = children

    @property
    def parent(self):
        return self._parent

    @parent.setter
    def parent(self, new_parent):
        if self._parent is not None:
            self._parent._children.remove(self)
        self._parent = new_parent
        if new_parent is not None:
            new_parent._children.append(self)

    @property
    def children(self):
        return tuple(self._children)

    @children.setter
    def children(self, new_children):
        for c in list(self._children):
            c._parent = None
        self._children = []
        for c in new_children:
            c.parent = self

    @property
    def is_root(self):
        return self._parent is None

    @property
    def is_leaf(self):
        return not self._children


class RenderTree:
    def __init__(self, node):
        self.node = node

    def __iter__(self):
        def walk(n, depth):
            yield ("  " * depth, "  " * depth, n)
            for c in n.children:
                yield from walk(c, depth + 1)
        return walk(self.node, 0)
