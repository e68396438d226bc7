#!/usr/bin/env python3
"""Headline benchmark: env-steps/s of the fused UR3e+2F85(+mug) step on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload rollout|reach|mug] [--envs-per-gpu E] [--dtype f32|f64]
    python bench.py --impl reference ...      # the CPU arm: oracle port of the reference step on all host cores

One "step" = one Env.step for every environment on every rank (one kernel launch per rank).  Prints ONE JSON line.
Workloads (BASELINE.json configs): rollout = config 3 (gymnasium_env/ur3e-v2 semantics on main.xml, U(action_space)
actions, auto-reset, 65536 envs per GPU) -- the configuration the 1e8 env-steps/s target is quoted on; reach = config 2
(ur3e_2f85.xml, pid_task_ctrl every mj_step, 4096 envs); mug = config 4 (main.xml scripted pick-and-lift, 16384 envs).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "env-steps/sec (UR3e+2F85, physics+controller)"
UNIT = "env-steps/s"


def parse():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=200)
    p.add_argument("--warmup", type=int, default=20)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--workload", default="rollout", choices=["rollout", "reach", "mug"])
    p.add_argument("--envs-per-gpu", type=int, default=0)
    p.add_argument("--dtype", default="f32", choices=["f32", "f64"])
    p.add_argument("--cpu-steps", type=int, default=0, help="env-steps per CPU worker for the cpu_baseline sample (0 = auto)")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-e2e", action="store_true")
    p.add_argument("--solver-iters", type=int, default=0, help="Newton iteration cap (0 = library default for the dtype)")
    p.add_argument("--no-l2-flush", action="store_true", help="profiling aid: skip the L2 flush so that ncu's dram counters show the kernel's own traffic")
    return p.parse_args()


WORKLOADS = {
    "rollout": dict(envs=65536, desc="config 3: gymnasium_env/ur3e-v2 rollout collection on main.xml, U(action_space) actions, auto-reset, frame_skip 2, dt 1 ms"),
    "reach": dict(envs=4096, desc="config 2: ur3e_2f85.xml task-space reach, pid_task_ctrl every mj_step, contact-free, frame_skip 1, dt 1 ms"),
    "mug": dict(envs=16384, desc="config 4: main.xml scripted pick-and-lift through the ur3e-v2 wrapper (approach + closing untimed, the timed region is the grasp / lift phase with gripper-mug contacts), frame_skip 2"),
}


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons during the timed region (B200_PROFILING.md clocks line)."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.stop_flag, self.rows = index, threading.Event(), []

    def run(self):
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self.stop_flag.wait(0.2)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows for i in range(4) if len(r) > 2 + i and r[2 + i].lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(self.rows)}


def cpu_leg(workload, steps_per_proc, procs=None):
    """Oracle port of the reference step on the host cores (bounded sample)."""
    from oracle import cpu_bench
    xml = os.path.join(ROOT, "ur3e_b200", "assets", "main.xml")
    # all three workloads' CPU legs run the v2 env step on main.xml, the reference's training env (train_rl.py:38-44);
    # the reach workload (no mug) is cheaper per step, so this is a conservative (slower) CPU number only for `reach`.
    rate, procs, total, wall = cpu_bench.run(xml, "v2", steps_per_proc, procs)
    return rate, procs, total, wall


def flops_per_env_step(model_name, frame_skip, nefc, ncon, it):
    """SURVEY 8(d) algorithmic FLOP formula (FMA = 2) with measured mean nefc / ncon / Newton iterations."""
    C = dict(A=dict(nb=7, nv=6, nM=21, S2=91, grip=0), B=dict(nb=20, nv=14, nM=81, S2=543, grip=1), C=dict(nb=21, nv=20, nM=102, S2=634, grip=1))[model_name]
    nb, nv, nM, S2 = C["nb"], C["nv"], C["nM"], C["S2"]
    nbb, nbp = (4, 5) if model_name == "C" else (0, 0)
    f = 190 * nb + 100 * nb + 20 * nv + 10 * nb + 40 * nv + 11 * nM + 2 * S2 + 190 * nb + 40 * nv + 10 * nv + 4 * nM
    f += 1500 * nbb + 100 * nbp + 1820 * C["grip"] + 800 * ncon + nefc * (4 * nM + 2 * nv) + it * nefc * 4 * nv + 2 * S2 + 4 * nM + 4 * nv
    return frame_skip * f + 600 + 300


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path.  mujoco/gymnasium are not installable here and the
    reference's meshes are absent, so this is the oracle port (kind 'port'), all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per = args.cpu_steps or 8000
    t_all = []
    total_steps = 0
    for _ in range(max(1, min(args.steps, 3))):    # bounded: a few samples, each ~ per x cores env-steps
        rate, procs, total, wall = cpu_leg(args.workload, per, cores)
        t_all.append(rate); total_steps += total
    value = sorted(t_all)[len(t_all) // 2]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "impl": "reference", "config": {"workload": WORKLOADS[args.workload]["desc"], "note": "one step here = one env-step of one CPU env; value aggregates all host cores"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d processes x %d ur3e-v2 env-steps (oracle float64 restatement; MuJoCo itself is not installable here), median of %d runs" % (cores, per, len(t_all))},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import numpy as np
    import torch
    import torch.distributed as dist
    import ur3e_b200._lib as lib
    from ur3e_b200 import presets
    from ur3e_b200.batch import SimBatch
    from ur3e_b200.model import Model, asset

    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")   # NCCL's own banner / debug lines must not share stdout with the JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dtype = torch.float32 if args.dtype == "f32" else torch.float64
    n = args.envs_per_gpu or WORKLOADS[args.workload]["envs"]

    if args.workload == "reach":
        model = Model(asset("ur3e_2f85.xml")); mname = "B"
        cfg = presets.make_config(model, dict(ctrl_mode=lib.CTRL_PID_TASK, obs_kind=lib.OBS_STATE, obs_dim=28, act_dim=7, frame_skip=1, gains=presets.GAINS_L_TASK,
                                             reset_key="down"), env_id_base=rank * n)
    else:
        model = Model(asset("main.xml")); mname = "C"
        xml, kw, _, _ = presets.ENV_SPECS["gymnasium_env/ur3e-v2"]
        cfg = presets.make_config(model, kw, auto_reset=1, env_id_base=rank * n, reset_noise=lib.NOISE_HIGH if args.workload == "rollout" else lib.NOISE_LOW,
                                  solver_iterations=args.solver_iters)
    batch = SimBatch(model, cfg, n, local, dtype)
    obs0 = batch.reset(seed=0).clone()
    ki = batch.kernel_info()

    # synthetic action streams, resident in HBM before the timed region
    g = torch.Generator(device=dev); g.manual_seed(1234 + rank)
    NBUF = 16
    if args.workload == "rollout":
        lo, hi = presets.action_bounds(model, "gymnasium_env/ur3e-v2")
        lo_t, hi_t = torch.tensor(lo, device=dev, dtype=dtype), torch.tensor(hi, device=dev, dtype=dtype)
        acts = [(lo_t + (hi_t - lo_t) * torch.rand(n, 4, device=dev, dtype=dtype, generator=g)).contiguous() for _ in range(NBUF)]
    elif args.workload == "reach":
        tcp = torch.tensor([0.29799994, 0.13349916, 0.1682003], device=dev, dtype=dtype)   # tcp at keyframe 'down' (assets/main.xml:415)
        acts = []
        for _ in range(NBUF):
            a = torch.zeros(n, 7, device=dev, dtype=dtype)
            a[:, :3] = tcp + (torch.rand(n, 3, device=dev, dtype=dtype, generator=g) - 0.5) * 0.2
            a[:, 3:6] = torch.tensor(presets.TOOL_ROTVEC, device=dev, dtype=dtype)
            acts.append(a.contiguous())
    else:
        acts = None   # scripted from the observation, see below

    mug0 = obs0[:, 3:6].clone()     # mug position at reset (the script tracks the live mug position in x, y)

    def mug_action(k):
        # scripted pick-and-lift (build_traj_l_pick_place logic, controller/build_traj.py:28-59): descend over the mug with the
        # gripper open (k < 90), close on it, then lift and lower slowly so that the pads stay in contact with the mug
        o = batch.obs
        a = torch.empty(n, 4, device=dev, dtype=dtype)
        a[:, 0:2] = o[:, 3:5]
        if k < 170:
            a[:, 2] = mug0[:, 2] + 0.02 + max(0.0, 0.1 - 0.002 * k)
        else:
            ph = (k - 170) % 400
            a[:, 2] = mug0[:, 2] + 0.02 + 0.0004 * (ph if ph < 200 else 400 - ph)
        a[:, 3] = 1.0 if k > 90 else 0.0
        return a

    flush = torch.empty(192 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def one_step(k):
        a = acts[k % NBUF] if acts is not None else mug_action(k)
        batch.step(a, want_final_obs=False)

    n_warm = max(args.warmup, 3) if args.workload != "mug" else max(args.warmup, 200)   # mug: the approach + closing phase is untimed
    for k in range(n_warm):
        one_step(k)
    batch.stats(reset=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler = ClockSampler(local); sampler.start()
    launches0 = batch.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    for k in range(args.steps):
        if not args.no_l2_flush:
            flush.zero_()                   # L2 flush between timed iterations (outside the per-step event pair)
        a = acts[k % NBUF] if acts is not None else mug_action(n_warm + k)
        ev[k][0].record()
        batch.step(a, want_final_obs=False)
        ev[k][1].record()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.stop_flag.set(); sampler.join(2)
    step_ms = [s.elapsed_time(e) for s, e in ev]
    total_ms = float(sum(step_ms))
    launches = batch.launch_count - launches0
    ki = batch.kernel_info()
    st = batch.stats_dict(reset=True)
    t = torch.tensor([total_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        sv = batch.stats(reset=False).clone(); sv[:] = torch.tensor([st[k] for k in lib.STAT_NAMES] + [0, 0], device=dev, dtype=torch.float64)
        dist.all_reduce(sv, op=dist.ReduceOp.SUM)      # the optional NCCL episode-stat all-reduce, off the step path
        st = {k: float(sv[i].item()) for i, k in enumerate(lib.STAT_NAMES)}
    total_ms = float(t.item())
    value = n * world * args.steps / (total_ms * 1e-3)

    # ---- e2e: the host-buffer C-ABI call, pinned host buffers, H2D + D2H inside the timed region
    e2e = None
    if not args.no_e2e:
        npdt = np.float32 if dtype == torch.float32 else np.float64
        hb = [torch.empty(n, batch.act_dim, dtype=dtype).pin_memory() for _ in range(4)]
        for i, h in enumerate(hb):
            h.copy_(acts[i % NBUF].cpu() if acts is not None else mug_action(n_warm + args.steps + i).cpu())
        h_obs = torch.empty(n, batch.obs_dim, dtype=dtype).pin_memory(); h_rew = torch.empty(n, dtype=dtype).pin_memory()
        h_te = torch.empty(n, dtype=torch.uint8).pin_memory(); h_tr = torch.empty(n, dtype=torch.uint8).pin_memory()
        ke = max(10, min(args.steps, 50))
        for k in range(3):
            batch.step_host(hb[k % 4], h_obs, h_rew, h_te, h_tr)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for k in range(ke):
            batch.step_host(hb[k % 4], h_obs, h_rew, h_te, h_tr)
        torch.cuda.synchronize()
        te = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        esz = 4 if dtype == torch.float32 else 8
        e2e = {"value": n * world * ke / float(te.item()), "unit": UNIT, "h2d_bytes_per_step": n * batch.act_dim * esz,
               "d2h_bytes_per_step": n * (batch.obs_dim * esz + esz + 2), "steps": ke}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant (only) kernel on the step path
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
    esz = 4 if dtype == torch.float32 else 8
    bytes_per_env = 2 * ki["state_bytes"] + batch.act_dim * esz + batch.obs_dim * esz + esz + 2
    kernel_ms = total_ms / max(launches, 1) if acts is not None else float(np.mean(step_ms))   # one launch per step
    sub = max(st["substeps"], 1.0)
    mean_nefc, mean_ncon, mean_it = st["nefc_sum"] / sub, st["ncon_sum"] / sub, st["solver_iter_sum"] / sub
    fs = cfg.frame_skip
    flops_env = flops_per_env_step(mname, fs, mean_nefc, mean_ncon, mean_it)
    props = torch.cuda.get_device_properties(dev)
    sm_mhz = float(peaks.get("sm_max_mhz", 1965.0))
    fp32_peak = props.multi_processor_count * 128 * 2 * sm_mhz * 1e6 / 1e12
    per_gpu_rate = n * args.steps / (total_ms * 1e-3)
    traffic = None
    try:   # dram__bytes_read + dram__bytes_write of the dominant kernel from the committed ncu capture, scaled per launch
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_traffic.json")))
        if args.workload == "rollout" and args.dtype == "f32":
            traffic = (tj["dram_bytes_read"] + tj["dram_bytes_write"]) * n / tj["envs"]
    except Exception:
        pass
    roof = {"bound": "hbm", "achieved": bytes_per_env * n / (kernel_ms * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s", "traffic": traffic,
            "peak_source": peak_src, "algorithmic_bytes_per_env_step": bytes_per_env,
            "note": "the path is FP32-pipe/latency bound, not HBM bound (SURVEY 8d): see roofline_fp32"}
    roof["frac"] = roof["achieved"] / roof["peak"]
    roof32 = {"bound": "fp32", "achieved": per_gpu_rate * flops_env / 1e12, "peak": fp32_peak, "unit": "TFLOP/s",
              "algorithmic_flops_per_env_step": flops_env, "mean_nefc": mean_nefc, "mean_ncon": mean_ncon, "mean_newton_iters": mean_it,
              "peak_source": "SMs x 128 lanes x 2 x sm_max_mhz"}
    roof32["frac"] = roof32["achieved"] / roof32["peak"]

    cpu = None
    if not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        per = args.cpu_steps or 80000
        rate, procs, total, wall = cpu_leg(args.workload, per, cores)
        cpu = {"value": rate, "unit": UNIT, "cores": procs, "kind": "port",
               "sample": "%d processes x %d ur3e-v2 env-steps of the float64 oracle restatement (%.1f s wall); MuJoCo 3.3.3 is not installable here" % (procs, per, wall)}

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": n_warm,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOADS[args.workload]["desc"], "envs_per_gpu": n, "total_envs": n * world, "frame_skip": fs, "substeps_per_s": value * fs,
                       "l2": "192 MiB buffer written between timed iterations (L2 flush), outside the per-step event pairs",
                       "parallelism": "independent env shards, one process per GPU, no step-path collective",
                       "kernel": ki, "episodes": st["episodes"], "unstable_resets": st["unstable_resets"], "overflow_steps": st["overflow_steps"]},
            "roofline": roof, "roofline_fp32": roof32, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": sampler.summary()}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
