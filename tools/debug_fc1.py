import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
from tests.util import batch_of, make_agent
from oracle import sac_oracle_np as O
name = sys.argv[1] if len(sys.argv) > 1 else "ckpt376_m1"
case = cases.UPDATE_CASES[name]
res = {}
for math in ("fp32", "tf32x3"):
    agent, st = make_agent(hw, case, math=math)
    b = batch_of(case, 0)
    if math == "fp32":
        ref_l, aux = O.update_parameters(st, b, return_aux=True)
    agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]), export_grads=True)
    res[math] = {n: agent.exported_grads(n) for n in ("q1", "q2", "policy")}
for net in ("q1", "q2", "policy"):
    for nm in res["fp32"][net]:
        a, c, r = res["fp32"][net][nm], res["tf32x3"][net][nm], aux[f"{net}_grads"][nm].reshape(res["fp32"][net][nm].shape)
        e1 = np.linalg.norm(a - r) / np.linalg.norm(r); e2 = np.linalg.norm(c - r) / np.linalg.norm(r)
        print(f"{net:7s} {nm:16s} fp32 {e1:.2e}  x3 {e2:.2e}")
        if e2 > 3e-4 and a.ndim == 2:
            d = np.abs(c - r)
            M, N = d.shape
            print("   tile map (128x64) of max|diff| / max|ref| :")
            for i in range(0, M, 128):
                print("   ", " ".join(f"{d[i:i+128, j:j+64].max() / np.abs(r).max():.1e}" for j in range(0, N, 64)))
            bad = np.argwhere(d > 0.2 * d.max())
            print("   worst entries rows:", sorted(set(bad[:, 0]))[:20], "cols:", sorted(set(bad[:, 1]))[:20])
