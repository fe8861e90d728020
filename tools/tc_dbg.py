import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import net_ref as nr
from connect4_b200.neural.model import ModelWrapper
g = np.load(os.path.join(ROOT, "tests/golden/net_outputs.npz"))
sd = nr.load_golden_state(os.path.join(ROOT, "tests/golden/example_net_state.npz"))
tc = ModelWrapper(state_dict=sd)
c0 = np.tile(g["c0"], 3)[:4096]; c1 = np.tile(g["c1"], 3)[:4096]
for i in range(3):
    tc.evaluate_bitboards(c0, c1)
torch.cuda.synchronize()
