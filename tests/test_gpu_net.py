"""CUDA network kernel vs the fp32 torch restatement / the reference's own outputs (tolerance 1e-2 absolute on
values and priors, the bound BASELINE.json's north_star states for the bf16 path)."""
import os

import numpy as np
import pytest

from conftest import GOLDEN, golden, random_positions

pytestmark = pytest.mark.gpu
TOL = 1e-2


def _golden_model():
    from oracle import net_ref as nr
    from connect4_b200.neural.model import ModelWrapper
    sd = nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz"))
    return ModelWrapper(state_dict=sd), sd


def test_example_net_vs_reference_outputs():
    """the reference's trained checkpoint on 1536 positions: outputs recorded from the reference's ModelWrapper"""
    g = golden("net_outputs.npz")
    model, _ = _golden_model()
    v, p = model.evaluate_bitboards(g["c0"], g["c1"])
    v, p = v.cpu().numpy(), p.cpu().numpy()
    dv, dp = np.abs(v - g["value"]).max(), np.abs(p - g["prior"]).max()
    print("example_net max|dvalue| %.2e max|dprior| %.2e" % (dv, dp))
    assert dv < TOL and dp < TOL
    assert np.abs(p.sum(axis=1) - 1).max() < 1e-5
    assert model.flops_per_position == 4740876.0


def test_model_wrapper_call_protocol():
    """model(Board) -> ((1,), (7,)); model([Board]) -> ((N,), (N,7)); TypeError otherwise (model.py:171-178)"""
    from connect4_b200.board import Board
    model, _ = _golden_model()
    v, p = model(Board())
    assert v.shape == (1,) and p.shape == (7,) and v.dtype == np.float32 and p.dtype == np.float32
    assert abs(float(v[0]) - 0.5381826) < TOL
    b2 = Board()
    b2.make_move(3)
    vs, ps = model([Board(), b2])
    assert vs.shape == (2,) and ps.shape == (2, 7)
    assert np.array_equal(vs[:1], v) and np.array_equal(ps[0], p)      # batch composition does not change a row
    with pytest.raises(TypeError):
        model("board")


def test_random_init_nets_vs_torch_fp32():
    """seeded random-init default net and the example_config net (64 filters / 6 residual / 6 fc)"""
    import torch
    from oracle import net_ref as nr
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.model import ModelWrapper
    g = golden("net_outputs.npz")
    torch.manual_seed(0)
    small = ModelWrapper(ModelConfig())
    v, p = small.evaluate_bitboards(g["c0"][:256], g["c1"][:256])
    assert np.abs(v.cpu().numpy() - g["rand_value"]).max() < TOL and np.abs(p.cpu().numpy() - g["rand_prior"]).max() < TOL
    torch.manual_seed(0)
    big = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
    v, p = big.evaluate_bitboards(g["c0"][:256], g["c1"][:256])
    dv, dp = np.abs(v.cpu().numpy() - g["big_value"]).max(), np.abs(p.cpu().numpy() - g["big_prior"]).max()
    print("64f/6r net max|dvalue| %.2e max|dprior| %.2e" % (dv, dp))
    assert dv < TOL and dp < TOL
    assert big.flops_per_position == 37342620.0


def test_8ply_shaped_set_67557_positions():
    """BASELINE.json configs[4]: 67,557 synthetic 8-ply positions, value 'RMSE' (mean MSE, stats.py:14-16) parity"""
    import random
    import torch
    from oracle import c4oracle as o
    from oracle import net_ref as nr
    rng = random.Random(8)
    pos = []
    while len(pos) < 67557:
        c0 = c1 = 0
        ok = True
        for _ in range(8):
            m = o.legal_mask(c0, c1)
            c0, c1, res = o.drop(c0, c1, rng.choice([c for c in range(7) if m >> c & 1]))
            if res != -1:
                ok = False
                break
        if ok:
            pos.append((c0, c1))
    a = np.array(pos, np.uint64)
    model, sd = _golden_model()
    v, p = model.evaluate_bitboards(a[:, 0].copy(), a[:, 1].copy())
    v, p = v.cpu().numpy(), p.cpu().numpy()
    torch.set_num_threads(max(1, os.cpu_count() or 1))
    rv, rp = nr.evaluate(sd, a[:, 0], a[:, 1])
    dv, dp = np.abs(v - rv), np.abs(p - rp)
    print("8ply set: value max %.2e mean %.2e | prior max %.2e mean %.2e" % (dv.max(), dv.mean(), dp.max(), dp.mean()))
    assert dv.max() < TOL and dp.max() < TOL
    labels = np.round(rv * 2) / 2                                 # synthetic {0, 0.5, 1} labels
    mse_ref, mse_gpu = float(np.mean((rv - labels) ** 2)), float(np.mean((v - labels) ** 2))
    assert abs(mse_ref - mse_gpu) < 1e-3


def test_device_count_argument_and_ragged_sizes():
    """c4_net_forward honours a device-side count and any batch size (empty, 1, not a multiple of the tile)"""
    import ctypes as C
    import torch
    from connect4_b200 import _lib
    g = golden("net_outputs.npz")
    model, _ = _golden_model()
    full_v, full_p = model.evaluate_bitboards(g["c0"], g["c1"])
    for n in (1, 7, 8, 9, 1183, 1185):
        v, p = model.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        assert torch.equal(v, full_v[:n]) and torch.equal(p, full_p[:n])
    t0 = torch.as_tensor(g["c0"].view(np.int64)).cuda()
    t1 = torch.as_tensor(g["c1"].view(np.int64)).cuda()
    out = torch.full((len(g["c0"]), 8), -7.0, dtype=torch.float32, device="cuda")
    cnt = torch.tensor([100], dtype=torch.int32, device="cuda")
    _lib.check(_lib.load().c4_net_forward(model.c4_net, _lib.ptr(t0), _lib.ptr(t1), len(g["c0"]), _lib.ptr(cnt),
                                          _lib.ptr(out), _lib.stream_ptr()))
    torch.cuda.synchronize()
    assert torch.equal(out[:100, 7], full_v[:100]) and (out[100:] == -7.0).all()
    v, p = model.evaluate_bitboards(np.zeros(0, np.uint64), np.zeros(0, np.uint64))
    assert v.numel() == 0


def test_bf16_operand_mode_is_selectable_and_less_accurate():
    """bf16 operands (north_star's nominal dtype) are kept as an option; on the trained checkpoint their worst-case
    value error exceeds the 1e-2 parity bound, which is why fp16 operands are the default"""
    from oracle import net_ref as nr
    from connect4_b200.neural.model import ModelWrapper
    g = golden("net_outputs.npz")
    sd = nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz"))
    v16, p16 = ModelWrapper(state_dict=sd, operand_dtype="fp16").evaluate_bitboards(g["c0"], g["c1"])
    vb, pb = ModelWrapper(state_dict=sd, operand_dtype="bf16").evaluate_bitboards(g["c0"], g["c1"])
    e16 = np.abs(v16.cpu().numpy() - g["value"])
    eb = np.abs(vb.cpu().numpy() - g["value"])
    print("fp16 max %.2e mean %.2e | bf16 max %.2e mean %.2e" % (e16.max(), e16.mean(), eb.max(), eb.mean()))
    assert e16.max() < TOL and eb.mean() < 5e-3 and eb.max() < 0.1 and e16.mean() < eb.mean()


def test_tcgen05_and_mma_sync_kernels_agree():
    """the default tcgen05/TMEM tower and the mma.sync tower are two independent implementations of the same math:
    both within tolerance of the reference, and within 2.5e-3 of each other (different fp32 summation order)"""
    import torch
    from oracle import net_ref as nr
    from connect4_b200.neural.model import ModelWrapper
    g = golden("net_outputs.npz")
    sd = nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz"))
    tc = ModelWrapper(state_dict=sd, kernel="auto")
    mma = ModelWrapper(state_dict=sd, kernel="mma")
    for n in (1, 15, 16, 17, 147, 148, 149, 1536):
        v1, p1 = tc.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        v2, p2 = mma.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        v1, p1, v2, p2 = v1.cpu().numpy(), p1.cpu().numpy(), v2.cpu().numpy(), p2.cpu().numpy()
        assert np.abs(v1 - g["value"][:n]).max() < TOL and np.abs(p1 - g["prior"][:n]).max() < TOL
        assert np.abs(v2 - g["value"][:n]).max() < TOL and np.abs(p2 - g["prior"][:n]).max() < TOL
        assert np.abs(v1 - v2).max() < 2.5e-3 and np.abs(p1 - p2).max() < 2.5e-3
    # a position's result does not depend on where in the batch (strip / tile / CTA) it is evaluated
    perm = np.random.RandomState(0).permutation(1536)
    v3, p3 = tc.evaluate_bitboards(g["c0"][perm], g["c1"][perm])
    v1, p1 = tc.evaluate_bitboards(g["c0"], g["c1"])
    assert torch.equal(v3, v1[torch.as_tensor(perm).cuda()]) and torch.equal(p3, p1[torch.as_tensor(perm).cuda()])


def test_tcgen05_kernel_for_64_filter_networks():
    """the example_config network (64 filters / 6 residual / 6 fc) runs on the tcgen05 kernel too (N = 192, K = 64 per tap
    row, 6-board strips): within tolerance of the reference's torch outputs, close to the mma.sync tower, and
    independent of the batch position -- for batch sizes around the strip (6) and grid (148) boundaries"""
    import torch
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.model import ModelWrapper
    g = golden("net_outputs.npz")
    cfg = ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6))
    torch.manual_seed(0)
    tc = ModelWrapper(cfg, kernel="auto")
    torch.manual_seed(0)
    mma = ModelWrapper(cfg, kernel="mma")
    nb = len(g["big_value"])
    for n in (1, 5, 6, 7, 147, 148, 149, 887, 888, 889, 1536):
        v1, p1 = tc.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        v2, p2 = mma.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        v1, p1, v2, p2 = v1.cpu().numpy(), p1.cpu().numpy(), v2.cpu().numpy(), p2.cpu().numpy()
        m = min(n, nb)
        assert np.abs(v1[:m] - g["big_value"][:m]).max() < TOL and np.abs(p1[:m] - g["big_prior"][:m]).max() < TOL
        assert np.abs(v1 - v2).max() < 5e-3 and np.abs(p1 - p2).max() < 5e-3, (n, np.abs(v1 - v2).max())
    perm = np.random.RandomState(1).permutation(1536)
    v3, p3 = tc.evaluate_bitboards(g["c0"][perm], g["c1"][perm])
    v1, p1 = tc.evaluate_bitboards(g["c0"], g["c1"])
    assert torch.equal(v3, v1[torch.as_tensor(perm).cuda()]) and torch.equal(p3, p1[torch.as_tensor(perm).cuda()])


@pytest.mark.parametrize("filters,residuals,fc", [(32, 1, 1), (32, 5, 2), (64, 2, 1), (64, 6, 6)])
def test_network_geometries_with_lively_weights(filters, residuals, fc):
    """other depths / widths than the two stock configurations, with weights and batch-norm statistics perturbed so that
    the activations are far from the near-constant outputs of a fresh initialisation: both kernels against the fp32 torch
    restatement of the same state dict"""
    import torch
    from oracle import net_ref as nr
    from connect4_b200.neural.config import ModelConfig, NetConfig
    from connect4_b200.neural.model import ModelWrapper, _init_state_dict
    g = golden("net_outputs.npz")
    torch.manual_seed(filters + residuals)
    sd = _init_state_dict(NetConfig(filters=filters, n_fc_layers=fc, n_residuals=residuals))
    gen = torch.Generator().manual_seed(7)
    for k, v in sd.items():
        if k.endswith("running_var"):
            v.copy_(0.5 + torch.rand(v.shape, generator=gen))
        elif k.endswith("running_mean"):
            v.copy_(0.2 * torch.randn(v.shape, generator=gen))
        elif "batch_norm" in k and k.endswith("weight") or (k.startswith("body.0.1") and k.endswith("weight")):
            v.copy_(0.6 + 0.8 * torch.rand(v.shape, generator=gen))
        elif "batch_norm" in k and k.endswith("bias") or (k.startswith("body.0.1") and k.endswith("bias")):
            v.copy_(0.2 * torch.randn(v.shape, generator=gen))
        elif "conv" in k and k.endswith("weight") or k == "body.0.0.weight":
            v.mul_(2.0)
        elif ".fc" in k and k.endswith("weight"):
            v.mul_(2.5)
    sdn = {k: v.numpy() for k, v in sd.items()}
    n = 600
    rv, rp = nr.evaluate(sdn, g["c0"][:n], g["c1"][:n])
    assert rv.std() > 0.003 and rp.std() > 0.01                      # the positions really are told apart
    from connect4_b200._lib import C4Error
    for kernel in ("auto", "mma"):
        cfg = ModelConfig(net_config=NetConfig(filters=filters, n_fc_layers=fc, n_residuals=residuals))
        if kernel == "mma" and filters == 32 and residuals > 4:      # its weights must all be resident in shared memory
            with pytest.raises(C4Error):
                ModelWrapper(cfg, state_dict=sd, kernel=kernel)
            continue
        m = ModelWrapper(cfg, state_dict=sd, kernel=kernel)
        v, p = m.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
        dv, dp = np.abs(v.cpu().numpy() - rv).max(), np.abs(p.cpu().numpy() - rp).max()
        print("%df/%dr/%dfc %s: max|dvalue| %.2e max|dprior| %.2e" % (filters, residuals, fc, kernel, dv, dp))
        assert dv < TOL and dp < TOL


def _scaled_trunk_state(c):
    """the reference's checkpoint with every trunk activation multiplied by c and the heads' 1x1 convs divided by c: the
    same function (LeakyReLU is positively homogeneous), but fp32 activations of the order of 4 * c"""
    import torch
    from oracle import net_ref as nr
    sd = {k: torch.as_tensor(np.array(v)).clone() for k, v in nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz")).items()}
    sd["body.0.1.weight"] *= c
    sd["body.0.1.bias"] *= c
    n_res = len({k.split(".")[2] for k in sd if k.startswith("body.1.")})
    for i in range(n_res):
        for j in (1, 2):
            sd["body.1.%d.batch_norm%d.running_mean" % (i, j)] *= c
            sd["body.1.%d.batch_norm%d.bias" % (i, j)] *= c
    sd["value_head.conv1.weight"] /= c
    sd["policy_head.conv1.weight"] /= c
    return sd


def test_fp16_operands_cannot_overflow_trunk_scale_chosen_at_creation(monkeypatch):
    """a network whose fp32 activations exceed 1e5 (fp16 tops out at 65504): c4_net_create calibrates a power-of-two trunk
    scale, the outputs stay within 1e-2 of the fp32 torch restatement; without the calibration the same network overflows
    and the engine call FAILS -- a non-finite answer never reaches a tree (reference: model.py:258-263 asserts)"""
    from connect4_b200._lib import C4Error
    from connect4_b200.engine import Engine
    from connect4_b200.mcts import MCTSConfig
    from connect4_b200.neural.game_pool import SelfPlayPool
    from connect4_b200.neural.model import ModelWrapper
    from oracle import net_ref as nr
    sd = _scaled_trunk_state(3.0e5)
    g = golden("net_outputs.npz")
    c0, c1 = g["c0"], g["c1"]
    rv, rp = nr.evaluate(sd, c0, c1)
    assert np.abs(rv - g["value"]).max() < 1e-4                      # same function as the checkpoint (fp32 torch)
    m = ModelWrapper(state_dict=sd)
    assert m.trunk_scale_log2 >= 8
    v, p = m.evaluate_bitboards(c0, c1)
    assert np.isfinite(v.cpu().numpy()).all()
    assert np.abs(v.cpu().numpy() - rv).max() < 1e-2 and np.abs(p.cpu().numpy() - rp).max() < 1e-2
    assert ModelWrapper(state_dict=nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz"))).trunk_scale_log2 == 0
    # the calibrated network plays
    pool = SelfPlayPool(m, MCTSConfig(24, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=8, seed=1)
    assert len(pool.generate_records(8)) >= 56
    pool.engine.close()
    # without calibration the operands overflow: loud failure on every engine path, never NaN into a tree
    # (trunk factor 2e5: the BN-folded stem weights, at most 0.25 in the checkpoint, still fit fp16 -- at 3e5 the weight
    #  guard of c4_net_create alone would already scale the trunk)
    monkeypatch.setenv("C4_NET_NO_CALIBRATION", "1")
    bad = ModelWrapper(state_dict=_scaled_trunk_state(2.0e5))
    assert bad.trunk_scale_log2 == 0
    assert not np.isfinite(bad.evaluate_bitboards(c0[:64], c1[:64])[0].cpu().numpy()).all()
    with pytest.raises(AssertionError):
        from connect4_b200.board import Board
        bad(Board())
    for engine in ("fused", "lockstep"):
        monkeypatch.setenv("C4_ENGINE", engine)
        pool = SelfPlayPool(bad, MCTSConfig(24, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=8, seed=1)
        with pytest.raises(C4Error, match="non-finite"):
            pool.generate_records(8)
        pool.engine.close()
        eng = Engine(8, MCTSConfig(24))
        eng.set_net(bad)
        eng.begin(c0[:8], c1[:8])
        with pytest.raises(C4Error, match="non-finite"):
            eng.run("net")
        eng.close()


def test_nan_weights_fail_loudly():
    """NaN in a head weight (invisible to the range calibration of the trunk): every answer is NaN -> error code"""
    import torch
    from connect4_b200._lib import C4Error
    from connect4_b200.mcts import MCTSConfig
    from connect4_b200.neural.game_pool import SelfPlayPool
    from connect4_b200.neural.model import ModelWrapper
    from oracle import net_ref as nr
    sd = {k: torch.as_tensor(np.array(v)).clone() for k, v in nr.load_golden_state(os.path.join(GOLDEN, "example_net_state.npz")).items()}
    sd["policy_head.fc1.weight"][3, 5] = float("nan")
    bad = ModelWrapper(state_dict=sd)
    pool = SelfPlayPool(bad, MCTSConfig(16, 19652, 1.25, 0.3, 0.25, 6), concurrent_games=4, seed=1)
    with pytest.raises(C4Error, match="non-finite"):
        pool.generate_records(4)
    pool.engine.close()
    sd["body.0.0.weight"][0, 0, 0, 0] = float("nan")                 # NaN in the trunk: refused at creation
    with pytest.raises(C4Error):
        ModelWrapper(state_dict=sd)
