"""ctypes binding of oracle/c4_oracle.c -- the CPU checker for the CUDA path.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
--impl reference legs.  The product package never imports this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "c4_oracle.c")
BUILD_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(BUILD_DIR, "libc4oracle.so")

NEED_EVAL, DONE = 1, 0
RES_NONE, RES_XWIN, RES_DRAW, RES_OWIN = -1, 0, 1, 2


def build(force=False):
    """gcc the C restatement (seconds). -ffp-contract=off keeps Python's two-rounding  a*b + c ."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(BUILD_DIR, exist_ok=True)
    subprocess.check_call(["gcc", "-O2", "-fPIC", "-shared", "-ffp-contract=off", "-o", LIB, SRC, "-lm"])
    return LIB


class Config(C.Structure):
    _fields_ = [("simulations", C.c_int), ("pb_c_base", C.c_double), ("pb_c_init", C.c_double),
                ("alpha", C.c_double), ("frac", C.c_double), ("num_sampling_moves", C.c_int)]


def make_config(simulations, pb_c_base=19652, pb_c_init=1.25, alpha=0.0, frac=0.0, num_sampling_moves=0):
    return Config(int(simulations), float(pb_c_base), float(pb_c_init), float(alpha), float(frac),
                  int(num_sampling_moves))


_lib = None
u64 = C.c_uint64
P = C.POINTER


def lib():
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    L.c4o_has_win.argtypes = [u64]
    L.c4o_age.argtypes = [u64, u64]
    L.c4o_legal_mask.argtypes = [u64, u64, C.c_int]
    L.c4o_drop.argtypes = [P(u64), P(u64), C.c_int]
    L.c4o_result_of.argtypes = [u64, u64]
    L.c4o_fliplr.argtypes = [u64]
    L.c4o_fliplr.restype = u64
    L.c4o_symmetrical.argtypes = [u64, u64]
    L.c4o_to_planes.argtypes = [u64, u64, C.c_void_p]
    L.c4o_from_plane.argtypes = [C.c_void_p]
    L.c4o_from_plane.restype = u64
    L.c4o_make_random_ips.argtypes = [C.c_int, C.c_void_p, C.c_int]
    L.c4o_evaluate_centre.argtypes = [u64, u64]
    L.c4o_evaluate_centre.restype = C.c_double
    L.c4o_tree_new.argtypes = [P(Config), u64, u64]
    L.c4o_tree_new.restype = C.c_void_p
    L.c4o_tree_free.argtypes = [C.c_void_p]
    L.c4o_tree_set_noise.argtypes = [C.c_void_p, C.c_void_p]
    L.c4o_search_start.argtypes = [C.c_void_p, P(u64), P(u64)]
    L.c4o_search_advance.argtypes = [C.c_void_p, P(u64), P(u64)]
    L.c4o_search_supply.argtypes = [C.c_void_p, C.c_double, C.c_void_p, C.c_void_p, P(u64), P(u64)]
    L.c4o_search_centre.argtypes = [C.c_void_p]
    L.c4o_best_move.argtypes = [C.c_void_p]
    L.c4o_sample_move.argtypes = [C.c_void_p, C.c_double]
    L.c4o_values_policy.argtypes = [C.c_void_p, C.c_void_p]
    L.c4o_visit_policy.argtypes = [C.c_void_p, C.c_void_p]
    L.c4o_root_children.argtypes = [C.c_void_p] + [C.c_void_p] * 4
    L.c4o_root_stats.argtypes = [C.c_void_p] + [C.c_void_p] * 5
    L.c4o_tree_size.argtypes = [C.c_void_p]
    L.c4o_tree_dump.argtypes = [C.c_void_p] + [C.c_void_p] * 7
    L.c4o_sweep_centre.argtypes = [P(Config), C.c_int] + [C.c_void_p] * 11
    L.c4o_selfplay_centre.argtypes = [P(Config), u64, u64] + [C.c_void_p] * 7 + [P(C.c_int)]
    _lib = L
    return L


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


# ------------------------------------------------------------------ board helpers (scalar)
def has_win(b): return bool(lib().c4o_has_win(int(b)))
def age(c0, c1): return lib().c4o_age(int(c0), int(c1))
def legal_mask(c0, c1, result=RES_NONE): return lib().c4o_legal_mask(int(c0), int(c1), int(result))
def result_of(c0, c1): return lib().c4o_result_of(int(c0), int(c1))
def fliplr(b): return int(lib().c4o_fliplr(int(b)))
def symmetrical(c0, c1): return bool(lib().c4o_symmetrical(int(c0), int(c1)))
def evaluate_centre(c0, c1): return float(lib().c4o_evaluate_centre(int(c0), int(c1)))


def drop(c0, c1, move):
    a, b = u64(int(c0)), u64(int(c1))
    res = lib().c4o_drop(C.byref(a), C.byref(b), int(move))
    return a.value, b.value, res


def to_planes(c0, c1):
    out = np.zeros((3, 6, 7), np.uint8)
    lib().c4o_to_planes(int(c0), int(c1), _p(out))
    return out


def from_plane(plane):
    a = np.ascontiguousarray(plane, dtype=np.uint8)
    return int(lib().c4o_from_plane(_p(a)))


def make_random_ips(plies, cap=4096):
    out = np.zeros((cap, 2), np.uint64)
    n = lib().c4o_make_random_ips(int(plies), _p(out), cap)
    assert n <= cap
    return out[:n]


# ------------------------------------------------------------------ search
class Tree:
    """One search tree (oinkoink/tree.py Tree + the loop of mcts.py search) with a stepping evaluator interface."""

    def __init__(self, cfg, c0, c1):
        self.cfg = cfg
        self.h = lib().c4o_tree_new(C.byref(cfg), int(c0), int(c1))
        self._a, self._b = u64(0), u64(0)

    def __del__(self):
        if getattr(self, "h", None):
            lib().c4o_tree_free(self.h)
            self.h = None

    def set_noise(self, gamma7):
        g = np.ascontiguousarray(gamma7, np.float64)
        lib().c4o_tree_set_noise(self.h, _p(g))

    def start(self):
        st = lib().c4o_search_start(self.h, C.byref(self._a), C.byref(self._b))
        return st, self._a.value, self._b.value

    def supply(self, value, prior):
        prior = np.ascontiguousarray(prior)
        if prior.dtype == np.float32:
            st = lib().c4o_search_supply(self.h, float(value), None, _p(prior), C.byref(self._a), C.byref(self._b))
        else:
            prior = prior.astype(np.float64)
            st = lib().c4o_search_supply(self.h, float(value), _p(prior), None, C.byref(self._a), C.byref(self._b))
        return st, self._a.value, self._b.value

    def search_centre(self):
        lib().c4o_search_centre(self.h)
        return self

    def search(self, evaluator):
        """evaluator(c0, c1) -> (value, prior ndarray[7] float64|float32)"""
        st, a, b = self.start()
        while st == NEED_EVAL:
            v, p = evaluator(a, b)
            st, a, b = self.supply(v, p)
        return self

    def best_move(self): return lib().c4o_best_move(self.h)
    def sample_move(self, u): return lib().c4o_sample_move(self.h, float(u))

    def values_policy(self):
        out = np.zeros(7, np.float64)
        lib().c4o_values_policy(self.h, _p(out))
        return out

    def visit_policy(self):
        out = np.zeros(7, np.float64)
        lib().c4o_visit_policy(self.h, _p(out))
        return out

    def root_children(self):
        v = np.zeros(7, np.int32); s = np.zeros(7, np.float64); r = np.zeros(7, np.int8); a = np.zeros(7, np.float64)
        lib().c4o_root_children(self.h, _p(v), _p(s), _p(r), _p(a))
        return v, s, r, a

    def root_stats(self):
        v = np.zeros(1, np.int32); s = np.zeros(1, np.float64); p = np.zeros(7, np.float64)
        n = np.zeros(1, np.int32); d = np.zeros(1, np.int32)
        lib().c4o_root_stats(self.h, _p(v), _p(s), _p(p), _p(n), _p(d))
        return int(v[0]), float(s[0]), p, int(n[0]), int(d[0])

    def dump(self):
        n = lib().c4o_tree_size(self.h)
        out = dict(parent=np.zeros(n, np.int32), name=np.zeros(n, np.int8), result=np.zeros(n, np.int8),
                   visits=np.zeros(n, np.int32), vsum=np.zeros(n, np.float64),
                   c0=np.zeros(n, np.uint64), c1=np.zeros(n, np.uint64))
        lib().c4o_tree_dump(self.h, *[_p(out[k]) for k in ("parent", "name", "result", "visits", "vsum", "c0", "c1")])
        return out


def sweep_centre(cfg, c0, c1):
    c0 = np.ascontiguousarray(c0, np.uint64)
    c1 = np.ascontiguousarray(c1, np.uint64)
    n = len(c0)
    out = dict(visits=np.zeros((n, 7), np.int32), vsum=np.zeros((n, 7), np.float64), cres=np.zeros((n, 7), np.int8),
               best=np.zeros(n, np.int8), best_value=np.zeros(n, np.float64), vpolicy=np.zeros((n, 7), np.float64),
               nodes=np.zeros(n, np.int32), root_visits=np.zeros(n, np.int32), root_vsum=np.zeros(n, np.float64))
    lib().c4o_sweep_centre(C.byref(cfg), n, _p(c0), _p(c1), *[_p(out[k]) for k in (
        "visits", "vsum", "cres", "best", "best_value", "vpolicy", "nodes", "root_visits", "root_vsum")])
    return out


def selfplay_centre(cfg, noise=None, uniform=None, start=(0, 0)):
    c0 = np.zeros(42, np.uint64); c1 = np.zeros(42, np.uint64); mv = np.zeros(42, np.int8)
    val = np.zeros(42, np.float64); pol = np.zeros((42, 7), np.float64)
    res = C.c_int(0)
    nz = None if noise is None else np.ascontiguousarray(noise, np.float64)
    un = None if uniform is None else np.ascontiguousarray(uniform, np.float64)
    if nz is not None and nz.shape[0] < 42:
        nz = np.concatenate([nz, np.ones((42 - nz.shape[0], 7))])
    if un is not None and un.shape[0] < 42:
        un = np.concatenate([un, np.zeros(42 - un.shape[0])])
    n = lib().c4o_selfplay_centre(C.byref(cfg), int(start[0]), int(start[1]), _p(nz), _p(un),
                                  _p(c0), _p(c1), _p(mv), _p(val), _p(pol), C.byref(res))
    return dict(n=n, c0=c0[:n], c1=c1[:n], moves=mv[:n], values=val[:n], priors=pol[:n], result=res.value)
