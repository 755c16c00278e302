"""Small end-to-end exercise for compute-sanitizer: tiny + C1 updates in both math modes and launch modes, PER sample/update, select_action."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import humanoid_walking_with_sac_b200 as hw
from tests.golden import cases
from tests.util import batch_of, make_agent

for name, math, launch in (("tiny_m2", "bf16x3", "staged"), ("tiny_m1", "fp32", "staged"), ("c1_bipedal_m1", "bf16x3", "staged"), ("tiny_m2", "bf16x3", "persistent")):
    case = cases.UPDATE_CASES[name]
    agent, _ = make_agent(hw, case, math=math, launch=launch, capacity=512, replay="per")
    b = batch_of(case, 0)
    agent.replay_buffer.push_many(b["s"], b["a"], b["r"], b["s2"], b["d"])
    out = agent.update_from_batch(b, eps=(b["eps_next"], b["eps_cur"]))
    out2 = agent.update_parameters(case["batch"])
    a = agent.select_action(b["s"][0])
    if launch == "staged" and math == "bf16x3":      # pipelined learner step (second stream) + exact-fallback-prone priorities
        agent2, _ = make_agent(hw, case, math=math, launch=launch, capacity=4096, replay="per", per_weighted_loss=True)
        rng = np.random.RandomState(3)
        n = 3000
        agent2.replay_buffer.push_many(rng.standard_normal((n, case["obs"])), rng.uniform(-0.4, 0.4, (n, case["act"])), rng.standard_normal(n),
                                       rng.standard_normal((n, case["obs"])), np.zeros(n))
        pri = np.zeros(4096, np.float32); pri[:n] = np.exp(3.0 * rng.standard_normal(n))
        agent2.replay_buffer.set_priorities(pri)
        for _ in range(4):
            agent2.learner_step(case["batch"])
        print("learner_step", agent2.learner_step(case["batch"], sync=True))
    print(name, math, launch, out, out2, a[:2])
print("SANITIZE_RUN_OK")
