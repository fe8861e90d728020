"""Evaluation pass: ValueStats / PriorStats / CombinedStats accounting (oinkoink/neural/stats.py) and
ModelWrapper.evaluate / evaluate_value_only (oinkoink/neural/pytorch/model.py:180-198,307-342).
Golden `eval_stats.npz`: the unmodified reference (example_net.pth, CPU fp32) on 6,000 labelled 8-ply-shaped positions,
with its per-position network outputs, so the accounting is checked exactly on the CPU and the CUDA pass within the
network tolerance on the GPU."""
import json
import os

import numpy as np
import pytest

from connect4_b200.neural.stats import CombinedStats, ValueStats

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _gold():
    return np.load(os.path.join(GOLD, "eval_stats.npz"))


def _mse(x, y):
    return float(np.mean((x.astype(np.float32) - y.astype(np.float32)) ** 2, dtype=np.float32))


def _bce(x, y):
    x = x.astype(np.float64)
    t = -(y * np.maximum(np.log(x), -100.0) + (1.0 - y) * np.maximum(np.log1p(-x), -100.0))
    return float(np.mean(t))


def test_stats_accounting_matches_the_reference_on_its_own_outputs():
    z = _gold()
    want = json.loads(str(z["combined"]))
    cs = CombinedStats()
    for i in range(0, len(z["values"]), 4096):                       # evaluate(batch_size=4096, shuffle=False)
        s = slice(i, i + 4096)
        cs.update(z["value_out"][s], z["values"][s], _mse(z["value_out"][s], z["values"][s]),
                  z["prior_out"][s], z["priors"][s], _bce(z["prior_out"][s], z["priors"][s]))
    got = cs.to_dict()
    assert list(got) == list(want)                                   # same keys, same order
    for k in ("Accuracy", "prior Accuracy", "Smallest", "Largest"):
        assert got[k] == want[k], k
    assert got["Average"] == pytest.approx(want["Average"], rel=1e-6)
    assert got["Average loss"] == pytest.approx(want["Average loss"], rel=1e-5)
    assert got["prior Average loss"] == pytest.approx(want["prior Average loss"], rel=1e-5)
    assert {str(k): list(v) for k, v in got["correct"].items()} == want["correct"]
    assert repr(cs) == str(z["combined_repr"])                       # printed form, 5 decimals
    vs = ValueStats()
    vs.update(z["value_out"], z["values"], _mse(z["value_out"], z["values"]))
    assert repr(vs) == str(z["value_only_repr"])


def test_categories_are_thirds_of_the_unit_interval():
    got = ValueStats.categorise_predictions(np.array([0.0, 0.3333, 0.3334, 0.6666, 0.6667, 0.99]))
    assert got.tolist() == [0.0, 0.0, 0.5, 0.5, 1.0, 1.0]


@pytest.mark.gpu
def test_cuda_evaluation_pass_reproduces_the_reference_stats():
    import torch
    from connect4_b200.board import BoardBatch
    from connect4_b200.neural.data import Connect4Dataset
    from connect4_b200.neural.model import ModelWrapper
    from oracle import net_ref as nr
    z = _gold()
    sd = nr.load_golden_state(os.path.join(GOLD, "example_net_state.npz"))
    model = ModelWrapper(state_dict=sd)
    planes = BoardBatch(z["c0"], z["c1"]).to_planes("float32").cpu()
    ds = Connect4Dataset(planes, torch.as_tensor(z["values"]), torch.as_tensor(z["priors"]))
    want = json.loads(str(z["combined"]))
    got = model.evaluate(ds, batch_size=4096, shuffle=False).to_dict()
    n = len(ds)
    # the network is within 1e-2 of the fp32 reference, so a handful of predictions next to a category border may move
    assert abs(got["Accuracy"] - want["Accuracy"]) * n <= 25
    assert abs(got["prior Accuracy"] - want["prior Accuracy"]) * n <= 25
    assert got["Average loss"] == pytest.approx(want["Average loss"], abs=2e-3)
    assert got["prior Average loss"] == pytest.approx(want["prior Average loss"], abs=2e-3)
    assert got["Average"] == pytest.approx(want["Average"], abs=1e-3)
    for k, (tot, cor) in want["correct"].items():
        assert got["correct"][float(k)][0] == tot and abs(got["correct"][float(k)][1] - cor) <= 25
    vo = model.evaluate_value_only(Connect4Dataset(planes, torch.as_tensor(z["values"]), None)).to_dict()
    assert vo["Average loss"] == pytest.approx(want["Average loss"], abs=2e-3)
    assert abs(vo["Accuracy"] - want["Accuracy"]) * n <= 25
