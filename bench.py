#!/usr/bin/env python
"""SAC learner hot-path benchmark (BASELINE.json metric: SAC updates/sec, Humanoid-v5 shape, B=256; PER samples/sec).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (N=1) = BASELINE.json configs[1]: Humanoid-v5 SAC update (obs 348, act 17, 3x512 networks_model2 MLPs,
batch 256) over a 1M-transition prioritized replay ring resident in HBM.  One STEP = what the trainer loop does
per environment step on the learner side: prioritized sample of 256 (index draw + IS weights + gather) -> one
full update_parameters (target, twin-critic, actor, temperature, Adam, Polyak) -> priority write-back.

  value : steps/s with every input resident in HBM (uniforms / eps drawn on device), CUDA events on the
          library's stream, no host sync inside the timed region.
  e2e   : steps/s through the public API (`replay_buffer.push` of one fresh transition + `update_parameters(256)`
          with host-drawn uniforms copied H2D and the three losses read back D2H every step).
  N>1   : independent replicas (one agent per GPU, no collective: the single-agent path does not shard), weak scaling.
  --impl reference : the CPU restatement of the reference (numpy/OpenBLAS update + C PER sampler) on the host cores.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

OBS, ACT, HID, NH, B = 348, 17, 512, 3, 256
CAPACITY = 1_000_000
FLOP_PER_UPDATE = 5.382e9          # fwd + necessary bwd, B=256, C2 shapes
WORKLOAD = "Humanoid-v5 SAC update obs348/act17/hidden512x3 (networks_model2) B=256 + 1M-transition prioritized replay"


def synth_transitions(n, seed):
    rng = np.random.RandomState(seed)
    s = rng.standard_normal((n, OBS)).astype(np.float32)
    a = rng.uniform(-0.4, 0.4, (n, ACT)).astype(np.float32)
    r = rng.standard_normal(n).astype(np.float32)
    s2 = rng.standard_normal((n, OBS)).astype(np.float32)
    d = (rng.uniform(size=n) < 0.01).astype(np.float32)
    return s, a, r, s2, d


def synth_priorities(n, seed):
    return (np.abs(np.random.RandomState(seed).standard_normal(n)) + 1e-6).astype(np.float32)


class ClockSampler:
    """SM clock / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe): NVML polled every ~4 ms from a
    thread (the timed region of a default run is ~70 ms, too short for `nvidia-smi -lms`), `nvidia-smi` as the fallback."""

    REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, gpu):
        self.gpu, self.sm, self.mx, self.reasons, self.proc, self.stop, self.thread, self.rows = gpu, [], [], set(), None, False, None, []
        self.nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nvml = pynvml
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[gpu]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else gpu
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx.append(float(pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)))
        except Exception:
            self.nvml = None

    def sample_now(self):
        """One sample from the calling thread (the bench calls it while the queued steps are still running on the GPU)."""
        n = self.nvml
        if not n:
            return
        try:
            self.sm.append(float(n.nvmlDeviceGetClockInfo(self.handle, n.NVML_CLOCK_SM)))
            mask = int(n.nvmlDeviceGetCurrentClocksThrottleReasons(self.handle))
            for bit, name in self.REASONS:
                if mask & bit:
                    self.reasons.add(name)
        except Exception:
            pass

    def _poll(self):
        while not self.stop:
            self.sample_now()
            time.sleep(0.004)

    def __enter__(self):
        if self.nvml:
            self.sample_now()                       # absorbs NVML's slow first queries before the timed region starts ...
            self.sm.clear(); self.reasons.clear()   # ... and is not a sample of it
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return self
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            r = [x.strip() for x in line.split(",")]
            if len(r) >= 7 and r[0].replace(".", "").isdigit():
                self.sm.append(float(r[0])); self.mx.append(float(r[1]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[3:7]):
                    if v.lower().startswith("active"):
                        self.reasons.add(name)

    def __exit__(self, *a):
        self.stop = True
        if self.thread:
            self.thread.join(timeout=1.0)
        if self.proc:
            self.proc.terminate()

    def summary(self):
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": max(self.mx) if self.mx else None,
                "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml" if self.nvml else "nvidia-smi"}


def measured_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            p = json.load(f)
        return p, "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------------------------
# reference arm: CPU restatement of the reference on the host cores (the oracle; bench is allowed to time it)
# ------------------------------------------------------------------------------------------------------------------
def cpu_reference(steps, warmup, n_per=CAPACITY, seed=0):
    from oracle import per_oracle as PO
    from oracle import sac_oracle_np as O
    try:    # give the CPU arm every host core (torchrun exports OMP_NUM_THREADS=1, which would handicap it)
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:
        pass
    st = O.make_state(OBS, ACT, HID, NH, seed=seed, head_scale=0.25)
    pri = synth_priorities(n_per, seed)
    pa = PO.pow_alpha(pri)
    s, a, r, s2, d = synth_transitions(8192, seed)      # the gather source: contents do not change the arithmetic
    rng = np.random.RandomState(seed + 1)

    def step(i):
        u = rng.random_sample(B)
        idx, w = PO.sample(pa, u, PO.beta(1 + i))                                  # replay_buffer.py:48-68 at N = 1M
        j = idx % 8192
        batch = dict(s=s[j], a=a[j], r=r[j], s2=s2[j], d=d[j],
                     eps_next=rng.standard_normal((B, ACT)).astype(np.float32), eps_cur=rng.standard_normal((B, ACT)).astype(np.float32))
        _, aux = O.update_parameters(st, batch, return_aux=True)                    # sac_imp.py:74-144
        td = np.abs(aux["td1"]).astype(np.float32)
        PO.update_priorities(pri, idx, td)                                          # replay_buffer.py:84-87
        pa[idx] = PO.pow_alpha(pri[idx])

    for i in range(warmup):
        step(i)
    t0 = time.perf_counter()
    for i in range(steps):
        step(warmup + i)
    dt = time.perf_counter() - t0
    try:
        from threadpoolctl import threadpool_info
        cores = max([p.get("num_threads", 1) for p in threadpool_info()] + [1])
    except Exception:
        cores = os.cpu_count()
    return steps / dt, dt / steps * 1e3, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    steps = max(3, min(args.steps, 60))
    warmup = max(1, min(args.warmup, 3))
    ups, ms, cores = cpu_reference(steps, warmup)
    sample = f"{steps} steps of the full workload (PER sample at N=1M + update + priority write-back), {warmup} warm-up"
    line = {"impl": "reference", "metric": "SAC updates/sec (Humanoid-v5 shape, B=256, 1M-transition PER)", "value": ups, "unit": "updates/s",
            "n_gpus": args.gpus, "steps": steps, "warmup": warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": {"workload": WORKLOAD, "host": "cpu"},
            "cpu_baseline": {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": ups, "unit": "updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


# ------------------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------------------
def build_agent(hw, device, launch, math, seed):
    import torch
    hw.use_networks("model2")
    torch.manual_seed(seed)
    agent = hw.SAC(OBS, ACT, hidden_dim=HID, device=f"cuda:{device}", replay="per", capacity=CAPACITY, max_batch=B,
                   math=math, launch=launch, seed=seed, per_weighted_loss=True)
    # shrink the policy heads like the parity cases do: keeps tanh out of saturation on N(0,1) observations
    with torch.no_grad():
        agent.policy.mean.weight.mul_(0.25)
        agent.policy.log_std.weight.mul_(0.25)
    chunk = 125_000
    s, a, r, s2, d = synth_transitions(chunk, seed)
    for _ in range(CAPACITY // chunk):
        agent.replay_buffer.push_many(s, a, r, s2, d)
    agent.replay_buffer.set_priorities(synth_priorities(CAPACITY, seed))
    return agent


def run_ours(args):
    import torch
    import torch.distributed as dist
    import humanoid_walking_with_sac_b200 as hw
    N = hw._native
    lib = N.lib()

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    backend = os.environ.get("SACB_BENCH_BACKEND", "nccl")      # "gloo": diagnosis only (timing barrier without NCCL)
    red_dev = f"cuda:{local}" if backend == "nccl" else "cpu"
    if world > 1:
        if backend == "nccl":
            dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
        else:
            dist.init_process_group(backend)
    agent = build_agent(hw, local, args.launch, args.math, seed=rank)
    h = agent._h

    def device_step():
        if args.no_pipeline:      # the three calls back to back on one stream
            N.check(lib.sacb_per_sample(h, 0, None, B, None, None, None, None, None, None, None))
            N.check(lib.sacb_update(h, B, None, None, None, None, N.USE_LAST_SAMPLE | N.NO_LOSS_READBACK))
            N.check(lib.sacb_per_update_from_td(h, 0, B))
        else:                     # same work, the write-back and the next sample run under the tail of the update (second stream)
            N.check(lib.sacb_per_step(h, B, None, N.NO_LOSS_READBACK))

    def barrier():
        agent.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(3, args.warmup)):
        device_step()
    barrier()
    st0 = agent.stats()["kernel_launches"]
    ms = ctypes.c_float()
    with ClockSampler(local) as clk:
        N.check(lib.sacb_timer_start(h))
        for _ in range(args.steps):
            device_step()
        # (NVML is only queried from the sampler's own thread: a query from THIS thread between the last launch and the stop event
        #  stalled the submission of the queued steps by 30-100 ms in one run out of four)
        N.check(lib.sacb_timer_stop(h, ctypes.byref(ms)))
        if clk.nvml and not clk.sm:
            clk.sample_now()      # region shorter than one NVML query: one sample right behind it
        if not clk.nvml:
            time.sleep(0.15)
    barrier()
    launches = agent.stats()["kernel_launches"] - st0
    if world > 1:
        print(f"[rank {rank}] device {local}: {ms.value / args.steps:.4f} ms/step", file=sys.stderr)
    t = torch.tensor([ms.value], device=red_dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total = float(t.item())
    value = world * args.steps / (ms_total * 1e-3)

    # ---- roofline of the dominant kernel (the update program) measured live: update-only replays on the same stream
    upd_ms = ctypes.c_float()
    N.check(lib.sacb_time_update(h, B, max(50, args.steps), ctypes.byref(upd_ms)))
    stage_us = (ctypes.c_float * 64)()
    n_st = lib.sacb_time_stages(h, B, stage_us, 64)
    per_ms = ctypes.c_float()
    N.check(lib.sacb_timer_start(h))
    for _ in range(args.steps):
        N.check(lib.sacb_per_sample(h, 0, None, B, None, None, None, None, None, None, None))
    N.check(lib.sacb_timer_stop(h, ctypes.byref(per_ms)))
    tp = torch.tensor([per_ms.value], device=red_dev)
    if world > 1:
        dist.all_reduce(tp, op=dist.ReduceOp.MAX)
    per_call_ms = float(tp.item()) / args.steps      # max over ranks; every rank samples from its own 1 M-row table
    peaks, peak_src = measured_peaks()
    achieved_tf = FLOP_PER_UPDATE / (upd_ms.value * 1e-3) / 1e12
    peak_tf = float(peaks.get("bf16_tflops_sustained", 1400.0))
    traffic = None      # DRAM bytes of one update program (26 stage kernels, warm caches) from the committed ncu capture
    try:
        with open(os.path.join(ROOT, "profiles", "r01_update_traffic.json")) as f:
            traffic = float(json.load(f)["update_dram_bytes"])
    except Exception:
        pass
    roofline = {"bound": "tensor", "achieved": achieved_tf, "peak": peak_tf, "unit": "TFLOP/s", "frac": achieved_tf / peak_tf, "traffic": traffic,
                "traffic_source": "profiles/r01_update_stages_ncu.csv: dram__bytes_read.sum + dram__bytes_write.sum summed over the stage kernels of one update (algorithmic bytes if streamed from HBM: 63.7e6; the state stays L2 resident)",
                "kernel": "sac_update_kernel (one fused update program: %d launches/step in '%s' mode)" % (1 if args.launch == "persistent" else n_st, args.launch),
                "ms_per_launch_sum": upd_ms.value, "peak_source": peak_src + " bf16 dense, sustained (each product costs 3 bf16 MMAs: algorithmic FLOPs are counted once)",
                "note": "single-agent B=256 is latency/occupancy bound (SURVEY 8d): 26 dependent stages of <=0.27 GFLOP, ~3 us of fixed cost each (profiles/r01_summary.md)",
                "stage_us": [round(float(stage_us[i]), 2) for i in range(max(0, min(n_st, 64)))],
                "per_sample": {"ms_per_call": per_call_ms, "samples_per_s": B / (per_call_ms * 1e-3),
                               "achieved_GBps": (3 * 4.0 * CAPACITY + B * 4 * (2 * OBS + ACT + 2)) / (per_call_ms * 1e-3) / 1e9,
                               "peak_GBps": float(peaks.get("hbm_gbs", 6650.0)), "algorithmic_bytes": "3 passes over p_alpha (4 B x N) + B rows"}}

    # ---- e2e through the public API: push one transition, update_parameters(256) with host uniforms, losses read back
    s1, a1, r1, s21, d1 = synth_transitions(args.steps + 8, 12345 + rank)
    row_bytes = int(lib.sacb_row_floats(h)) * 4
    for i in range(3):
        agent.replay_buffer.push(s1[i], a1[i], r1[i], s21[i], bool(d1[i]))
        agent.update_parameters(B)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        agent.replay_buffer.push(s1[i], a1[i], float(r1[i]), s21[i], bool(d1[i]))
        out = agent.update_parameters(B)
    agent.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], device=red_dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e = world * args.steps / float(te.item())

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        ups, cms, cores = cpu_reference(steps=40, warmup=2)
        cpu = {"value": ups, "unit": "updates/s", "cores": cores, "kind": "port",
               "sample": "40 steps of the full workload on the host (numpy/OpenBLAS restatement of sac_imp.py:74-144 + C restatement of replay_buffer.py:48-87 at N=1M)"}

    if rank == 0:
        line = {"metric": "SAC updates/sec (Humanoid-v5 shape, B=256, 1M-transition PER)", "value": value, "unit": "updates/s", "n_gpus": world,
                "steps": args.steps, "warmup": max(3, args.warmup), "ms_per_step": ms_total / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": {"bf16x3": "f32 (bf16 hi/lo operand pairs, 3 tcgen05 MMAs per product, f32 accumulate in TMEM; f32 master weights / Adam)", "fp32": "f32 (FFMA on the bf16-pair operands)"}[args.math],
                "data": "synthetic",
                "config": {"workload": WORKLOAD, "launch": args.launch, "math": args.math, "replicas": world,
                           "step": "sample(256) -> update -> priority write-back, " + ("sequential on one stream" if args.no_pipeline else
                                   "software-pipelined: write-back and the next sample run on a second stream under the tail of the update (sacb_per_step; bitwise equal to the sequential order)"),
                           "l2_policy": "inputs larger than L2: 1M-row ring (2.9 GB) + 4 MB priority table re-read every step; weights/Adam state (63 MB) stay L2 resident by design"},
                "e2e": {"value": e2e, "unit": "updates/s", "h2d_bytes_per_step": row_bytes + 8 * B, "d2h_bytes_per_step": 12},
                "gpu_launches": int(launches), "clocks": clk.summary(), "roofline": roofline, "cpu_baseline": cpu,
                "per_samples_per_s": world * B / (per_call_ms * 1e-3),      # whole job: `world` independent prioritized buffers
                "last_losses": out}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--launch", default=os.environ.get("SACB_LAUNCH", "staged"), choices=["staged", "persistent"])
    ap.add_argument("--math", default=os.environ.get("SACB_MATH", "bf16x3"), choices=["bf16x3", "fp32"])
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="value: run sample / update / write-back sequentially on one stream")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
