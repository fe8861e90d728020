"""fp32 torch restatement of the reference value/policy network (oinkoink/neural/pytorch/model.py:20-134).

TEST INFRASTRUCTURE ONLY (the floating-point checker for the CUDA net kernel; see oracle/c4_oracle.c header).
Parity status: PINNED -- tests/test_oracle_golden.py checks `forward_state` on the reference's own checkpoint
(tests/golden/example_net_state.npz = net_state_dict of oinkoink/data/example_net.pth) against the outputs of the
reference's ModelWrapper stored in tests/golden/net_outputs.npz, and `RefNet`'s seeded random init against the
reference's `Net` (same construction order => same draws from torch.manual_seed).

Written functionally (F.conv2d / F.batch_norm on a state dict) rather than as the reference's module tree.
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

LEAKY = 0.01   # nn.LeakyReLU() default slope (model.py:31,43,68,104)
BN_EPS = 1e-5  # nn.BatchNorm2d default eps


def planes_from_bitboards(c0, c1):
    """oinkoink/board.py:64-82,147-154 Board.to_array, vectorised: uint64[n] x2 -> float32 [n,3,6,7] (row 0 = top)."""
    c0 = np.asarray(c0, np.uint64)
    c1 = np.asarray(c1, np.uint64)
    n = c0.shape[0]
    out = np.zeros((n, 3, 6, 7), np.float32)
    occ = c0 | c1
    age = np.zeros(n, np.int64)
    for r in range(6):
        for c in range(7):
            bit = np.uint64(7 * c + (5 - r))
            out[:, 1, r, c] = ((c0 >> bit) & np.uint64(1)).astype(np.float32)
            out[:, 2, r, c] = ((c1 >> bit) & np.uint64(1)).astype(np.float32)
            age += ((occ >> bit) & np.uint64(1)).astype(np.int64)
    out[:, 0] = (age % 2 == 0).astype(np.float32)[:, None, None]
    return out


def _bn(x, sd, prefix):
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"],
                        sd[prefix + ".weight"], sd[prefix + ".bias"], training=False, eps=BN_EPS)


def forward_state(sd, x):
    """Net.forward in eval mode (model.py:130-134) on a reference-layout state dict of float32 tensors."""
    sd = {k: (torch.as_tensor(v) if not torch.is_tensor(v) else v) for k, v in sd.items()}
    x = torch.as_tensor(x, dtype=torch.float32)
    n_res = len({k.split(".")[2] for k in sd if k.startswith("body.1.")})
    n_fc = len({k.split(".")[2] for k in sd if k.startswith("value_head.fcN.")})
    # stem: conv3x3(no bias) + BN + LeakyReLU  (model.py:20-31)
    h = F.leaky_relu(_bn(F.conv2d(x, sd["body.0.0.weight"], padding=1), sd, "body.0.1"), LEAKY)
    # residual tower (model.py:45-55)
    for i in range(n_res):
        p = "body.1.%d." % i
        r = h
        o = F.leaky_relu(_bn(F.conv2d(h, sd[p + "conv1.weight"], padding=1), sd, p + "batch_norm1"), LEAKY)
        o = _bn(F.conv2d(o, sd[p + "conv2.weight"], padding=1), sd, p + "batch_norm2")
        h = F.leaky_relu(o + r, LEAKY)
    # value head (model.py:77-91): 1x1 conv -> BN -> LeakyReLU -> n_fc x Linear(42,42) (no activation between)
    # -> LeakyReLU -> Linear(42,1) -> tanh -> (x + w1) * w2
    v = F.leaky_relu(_bn(F.conv2d(h, sd["value_head.conv1.weight"], sd["value_head.conv1.bias"]),
                         sd, "value_head.batch_norm"), LEAKY)
    v = v.reshape(v.shape[0], -1)
    for j in range(n_fc):
        v = F.linear(v, sd["value_head.fcN.%d.weight" % j], sd["value_head.fcN.%d.bias" % j])
    v = F.leaky_relu(v, LEAKY)
    v = torch.tanh(F.linear(v, sd["value_head.fc1.weight"], sd["value_head.fc1.bias"]))
    v = ((v + sd["value_head.w1"]) * sd["value_head.w2"]).reshape(-1)
    # policy head (model.py:107-117): 1x1 conv (2 ch) -> BN -> LeakyReLU -> flatten channel-major -> Linear(84,7) -> softmax
    p = F.leaky_relu(_bn(F.conv2d(h, sd["policy_head.conv1.weight"], sd["policy_head.conv1.bias"]),
                         sd, "policy_head.batch_norm"), LEAKY)
    p = F.linear(p.reshape(p.shape[0], -1), sd["policy_head.fc1.weight"], sd["policy_head.fc1.bias"])
    p = torch.softmax(p, dim=1)
    return v, p


class RefNet(nn.Module):
    """Parameter container with the reference's construction order and state-dict key names
    (model.py:120-128), used only to reproduce its seeded random initialisation."""

    def __init__(self, channels=3, filters=32, n_fc_layers=4, n_residuals=3):
        super().__init__()

        class Res(nn.Module):
            def __init__(s, f):
                super().__init__()
                s.conv1 = nn.Conv2d(f, f, 3, padding=1, bias=False)
                s.conv2 = nn.Conv2d(f, f, 3, padding=1, bias=False)
                s.batch_norm1 = nn.BatchNorm2d(f)
                s.batch_norm2 = nn.BatchNorm2d(f)

        class VH(nn.Module):
            def __init__(s, f, n):
                super().__init__()
                s.conv1 = nn.Conv2d(f, 1, 1)
                s.batch_norm = nn.BatchNorm2d(1)
                s.fcN = nn.Sequential(*[nn.Linear(42, 42) for _ in range(n)])
                s.fc1 = nn.Linear(42, 1)
                s.w1 = nn.Parameter(torch.tensor(1.0), requires_grad=False)
                s.w2 = nn.Parameter(torch.tensor(0.5), requires_grad=False)

        class PH(nn.Module):
            def __init__(s, f):
                super().__init__()
                s.conv1 = nn.Conv2d(f, 2, 1)
                s.batch_norm = nn.BatchNorm2d(2)
                s.fc1 = nn.Linear(84, 7)

        self.body = nn.Sequential(
            nn.Sequential(nn.Conv2d(channels, filters, 3, padding=1, bias=False), nn.BatchNorm2d(filters)),
            nn.Sequential(*[Res(filters) for _ in range(n_residuals)]))
        self.value_head = VH(filters, n_fc_layers)
        self.policy_head = PH(filters)

    def forward(self, x):
        return forward_state(self.state_dict(), x)


def random_state(seed=0, filters=32, n_fc_layers=4, n_residuals=3):
    """state dict of a freshly constructed reference Net under torch.manual_seed(seed) (numpy float32 arrays)."""
    torch.manual_seed(seed)
    net = RefNet(3, filters, n_fc_layers, n_residuals)
    return {k: v.detach().numpy().copy() for k, v in net.state_dict().items()}


def load_golden_state(path):
    z = np.load(path)
    return {k: z[k] for k in z.files}


def evaluate(sd, c0, c1, batch=4096):
    """(values float32[n], priors float32[n,7]) like ModelWrapper._call_list (model.py:269-282)."""
    vs, ps = [], []
    with torch.no_grad():
        for i in range(0, len(c0), batch):
            v, p = forward_state(sd, planes_from_bitboards(c0[i:i + batch], c1[i:i + batch]))
            vs.append(v.numpy())
            ps.append(p.numpy())
    return np.concatenate(vs), np.concatenate(ps)
