"""per-role cycle accounting of the tcgen05 network kernel: build with tools/build_variant.sh NAME -DC4_TC_PROFILE, run with
C4_LIB=.../libc4b200_NAME.so C4_TC_DEBUG=1 python tools/tc_dbg.py [--net64]"""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from connect4_b200.neural.model import ModelWrapper
g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
if "--net64" in sys.argv:
    from connect4_b200.neural.config import ModelConfig, NetConfig
    torch.manual_seed(0)
    tc = ModelWrapper(ModelConfig(net_config=NetConfig(filters=64, n_fc_layers=6, n_residuals=6)))
else:
    _z = np.load(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
    tc = ModelWrapper(state_dict={k: _z[k] for k in _z.files})
c0 = np.tile(g["c0"], 3)[:4096]; c1 = np.tile(g["c1"], 3)[:4096]
for i in range(3):
    tc.evaluate_bitboards(c0, c1)
torch.cuda.synchronize()
