"""Small driver for ncu: a few update steps of the C2 (Humanoid, 3x512) shape.  usage: profile_update.py [math] [launch] [steps]"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
from tests.util import batch_of, make_agent

math = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
launch = sys.argv[2] if len(sys.argv) > 2 else "staged"
steps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
case = cases.UPDATE_CASES["c2_humanoid_m2"]
agent, st = make_agent(hw, case, math=math, launch=launch)
b = batch_of(case, 0)
for i in range(steps):
    out = agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
print(out)
