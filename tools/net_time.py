"""network kernel timing at the batch sizes the self-play path produces (CUDA events, resident inputs)"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from connect4_b200.neural.model import ModelWrapper
g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
z = np.load(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
if "--net64" in sys.argv:                       # the reference's example_config network, random init (torch.manual_seed(0))
    from connect4_b200.neural.config import ModelConfig, NetConfig
    torch.manual_seed(0)
    m = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
    ref_v, ref_p = g["big_value"], g["big_prior"]
    g = {"c0": g["c0"][:len(ref_v)], "c1": g["c1"][:len(ref_v)]}
else:
    m = ModelWrapper(state_dict={k: z[k] for k in z.files})
    ref_v, ref_p = g["value"], g["prior"]
v, p = m.evaluate_bitboards(g["c0"], g["c1"])
print("max |dv| %.2e  max |dp| %.2e over %d golden positions" % (np.abs(v.cpu().numpy() - ref_v).max(), np.abs(p.cpu().numpy() - ref_p).max(), len(ref_v)))
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for n in (148, 1000, 1900, 2300, 3700, 4096, 8192, 67557):
    c0 = torch.as_tensor(np.resize(g["c0"], n).view(np.int64)).cuda(); c1 = torch.as_tensor(np.resize(g["c1"], n).view(np.int64)).cuda()
    from connect4_b200 import _lib
    L = _lib.load(); out = torch.empty((n, 8), dtype=torch.float32, device="cuda"); st = _lib.stream_ptr()
    call = lambda: L.c4_net_forward(m.c4_net, _lib.ptr(c0), _lib.ptr(c1), n, None, _lib.ptr(out), st)
    for _ in range(5): call()
    torch.cuda.synchronize(); e0.record()
    for _ in range(200): call()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 200 * 1e3
    print("n %6d  %.1f us  %.0f TFLOP/s" % (n, us, n * m.flops_per_position / us / 1e6))
