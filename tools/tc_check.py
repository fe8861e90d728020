"""quick check of the tcgen05 network kernel against the mma.sync kernel and the fp32 torch restatement + timing"""
import os, sys, time
import numpy as np
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from connect4_b200.neural.model import ModelWrapper

g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
_z = np.load(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
sd = {k: _z[k] for k in _z.files}
tc = ModelWrapper(state_dict=sd, kernel="auto")
mma = ModelWrapper(state_dict=sd, kernel="mma")
for n in (1, 2, 16, 17, 100, 148, 1536):
    v, p = tc.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
    torch.cuda.synchronize()
    v2, p2 = mma.evaluate_bitboards(g["c0"][:n], g["c1"][:n])
    v, p, v2, p2 = v.cpu().numpy(), p.cpu().numpy(), v2.cpu().numpy(), p2.cpu().numpy()
    print("n=%5d  tc vs ref: dv %.2e dp %.2e | mma vs ref: dv %.2e | tc vs mma: dv %.2e" % (
        n, np.abs(v - g["value"][:n]).max(), np.abs(p - g["prior"][:n]).max(), np.abs(v2 - g["value"][:n]).max(),
        np.abs(v - v2).max()), flush=True)
reps = 4096 // 1536 + 1
c0 = np.tile(g["c0"], reps)[:4096]; c1 = np.tile(g["c1"], reps)[:4096]
from connect4_b200.engine import _u64_tensor
from connect4_b200 import _lib
t0, t1 = _u64_tensor(c0), _u64_tensor(c1)
out = torch.empty((4096, 8), dtype=torch.float32, device="cuda")
for name, m in (("tc", tc), ("mma", mma)):
    for n in (4096, 3552, 2048):
        for _ in range(3):
            _lib.check(_lib.load().c4_net_forward(m.c4_net, _lib.ptr(t0), _lib.ptr(t1), n, None, _lib.ptr(out), None))
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            _lib.check(_lib.load().c4_net_forward(m.c4_net, _lib.ptr(t0), _lib.ptr(t1), n, None, _lib.ptr(out), None))
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        print("%s n=%d: %.1f us/launch, %.0f TFLOP/s" % (name, n, ms * 1e3, n * m.flops_per_position / (ms * 1e-3) / 1e12), flush=True)
