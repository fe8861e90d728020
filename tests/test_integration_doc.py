"""INTEGRATION.md is executable: the ctypes stub a maintainer of the reference would add is taken verbatim from the
document's code blocks and run against the library -- network call, batched search and generation -- and its results are
checked against the reference's goldens."""
import os
import re
import types

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, golden

pytestmark = pytest.mark.gpu


class _Board():                               # the only thing the stub needs from an oinkoink Board: .color
    def __init__(self, c0, c1):
        self.color = np.array([c0, c1], dtype=np.uint64).view(np.int64)


def _stub_namespace():
    from connect4_b200 import _build
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, re.S)
    assert len(blocks) >= 4
    ns = {}
    for b in blocks:
        exec(compile(b.replace('C.CDLL("libc4b200.so")', 'C.CDLL(%r)' % _build.LIB), "INTEGRATION.md", "exec"), ns)
    return ns


def test_the_documented_stub_runs_and_agrees_with_the_goldens():
    import torch
    ns = _stub_namespace()
    # 2a: the network call
    z = np.load(os.path.join(GOLDEN, "example_net_state.npz"))
    fake_net = types.SimpleNamespace(state_dict=lambda: {k: torch.as_tensor(z[k]) for k in z.files})
    h = ns["make_net"](fake_net)
    g = golden("net_outputs.npz")
    boards = [_Board(int(a), int(b)) for a, b in zip(g["c0"][:64], g["c1"][:64])]
    values, priors = ns["call_list"](h, boards)
    assert np.abs(values - g["value"][:64]).max() < 1e-2 and np.abs(priors - g["prior"][:64]).max() < 1e-2
    # 2b: batched deterministic searches == the reference's searches
    m = golden("mcts_sweep_800.npz")
    boards = [_Board(int(a), int(b)) for a, b in zip(m["c0"][:32], m["c1"][:32])]
    cfg = types.SimpleNamespace(simulations=800, pb_c_base=19652, pb_c_init=1.25)
    visits, policy, best = ns["search_many"](boards, cfg)
    assert (visits == m["visits"][:32]).all() and (best == m["best"][:32]).all()
    assert (policy.view(np.uint64) == m["vpolicy"][:32].view(np.uint64)).all()
    # 2c: a generation straight to the data.pth tensors
    az = types.SimpleNamespace(simulations=32, pb_c_base=19652, pb_c_init=1.25, root_dirichlet_alpha=0.3,
                               root_exploration_fraction=0.25, num_sampling_moves=6)
    b, v, p = ns["generate_games"](h, az, 12, concurrent=8, seed=3)
    assert b.shape[1:] == (3, 6, 7) and len(b) == len(v) == len(p) and len(b) % 2 == 0 and len(b) >= 2 * 7 * 12
    assert torch.equal(b[len(b) // 2:], torch.flip(b[:len(b) // 2], dims=[3]))
    assert set(np.unique(v.numpy()).tolist()) <= {0.0, 0.5, 1.0}
