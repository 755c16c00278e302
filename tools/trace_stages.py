import sys, os, ctypes
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if not os.environ.get("SACB_NOTRACE"):
    os.environ["SACB_TRACE"] = "1"
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
from tests.util import batch_of, make_agent
math = sys.argv[1] if len(sys.argv) > 1 else "bf16x3"
case = cases.UPDATE_CASES["c2_humanoid_m2"]
agent, st = make_agent(hw, case, math=math)
b = batch_of(case, 0)
agent.update_from_batch(b)
us = (ctypes.c_float * 64)()
n = hw._native.lib().sacb_time_stages(agent._h, 256, us, 64)
print([round(us[i], 1) for i in range(n)])
