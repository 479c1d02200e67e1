#!/usr/bin/env python
"""bench.py -- headline benchmark of the batched time-stepping contact hot path.

A "step" is one TimeSteppingSimulator::step(dt = 1 ms) of every env of the workload (BASELINE.json configs[1]:
65,536 randomized sitting-box / bouncing-ball envs per GPU, SURVEY.md 8(d) case 2).  Envs are independent, so
N GPUs each own their own 65,536 envs (weak scaling, no collective on the step path); one NCCL all-reduce (max of
the elapsed time, sum of the counters) closes the run.

  python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
  python bench.py --impl reference --gpus N --steps K ...   # the CPU restatement of the reference on the host cores

Prints ONE JSON line (rank 0).  Timing: CUDA events on the launching stream around each step kernel, L2 flushed
between steps (a 256 MiB write, outside the events), max over ranks.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT]

ENVS_PER_GPU = 65536
DT = 1e-3
WORKLOAD = "sitting-box/bouncing-ball batch: 65,536 randomized envs per GPU (BASELINE configs[1]; SURVEY 8d case 2)"
BYTES_PER_ENV_STEP = 208.0        # 2 * 8 B * (7 q + 6 v) for the one moving body


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    hbm, src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(p):
        try:
            hbm, src = float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    fp64, fsrc = 36.07, "measured r01 (profiles/r01_fp64_peak.json)"
    q = os.path.join(ROOT, "profiles", "r01_fp64_peak.json")
    if os.path.exists(q):
        try:
            fp64 = float(json.load(open(q))["fp64_dfma_tflops"])
        except Exception:
            pass
    return hbm, src, fp64, fsrc


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.idx, self.lines, self.p = gpu_index, [], None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                       "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=lambda: [self.lines.append(l) for l in self.p.stdout], daemon=True).start()
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        sm, mx, reasons = [], None, set()
        for l in self.lines:
            f = [x.strip() for x in l.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx = float(f[2])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(sm)}


def _cpu_baseline(scene, q, v, n_envs, n_steps, threads):
    """The oracle (CPU restatement of the reference; the reference itself cannot be built here) on the host cores."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    import copy
    sc = copy.copy(scene)
    sc.q, sc.v = q, v
    batch = O.OracleBatch(sc, 0, n_envs)
    t0 = time.perf_counter()
    c = batch.run(DT, n_steps, threads=threads)
    el = time.perf_counter() - t0
    return c["env_steps"] / el, c["lcp_solves"] / el, el, c


def run_reference(args, rank, world):
    """--impl reference: the reference's own CPU implementation of the path, restated (oracle/, kind "port"),
    all host threads, each step = one step of a bounded 4,096-env sample of the same seeded workload."""
    if rank != 0:
        return
    from moby_b200 import scenes
    sample = 4096
    cores = os.cpu_count() or 1
    scene = scenes.small_lcp_batch(sample, seed=0xB200)
    if args.min_step == "default":
        scene.min_step_size_env = None
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_api as O
    O.build()
    batch = O.OracleBatch(scene, 0, sample)
    batch.run(DT, args.preroll, threads=cores)
    for _ in range(args.warmup):
        batch.run(DT, 1, threads=cores)
    c0 = batch.run(DT, 0, threads=1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        c1 = batch.run(DT, 1, threads=cores)
    el = time.perf_counter() - t0
    val = sample * args.steps / el
    out = {
        "impl": "reference", "metric": "env_steps_per_s", "value": val, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * el / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "dt": DT, "preroll_steps": args.preroll, "seed": "0xB200", "min_step_size": args.min_step},
        "lcp_solves_per_s": (c1["lcp_solves"] - c0["lcp_solves"]) / el,
        "cpu_baseline": {"value": val, "unit": "env-steps/s", "cores": cores, "kind": "port",
                         "sample": f"first {sample} envs of the seeded batch, one step per timed step, {cores} host threads; "
                                   "the reference cannot be compiled here (Ravelin/Boost/qhull/libxml2 absent), so this is the "
                                   "oracle/ restatement"},
        "e2e": {"value": val, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--envs-per-gpu", type=int, default=ENVS_PER_GPU)
    ap.add_argument("--preroll", type=int, default=300, help="untimed steps before warm-up so contacts are active")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--min-step", default="scene", choices=["scene", "default"],
                    help="scene: min-step-size of the source scenes (test/box.xml: 1e-3, bouncing-ball.xml: sqrt(eps)); "
                         "default: sqrt(eps) everywhere (TimeSteppingSimulator.cpp:48)")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from moby_b200 import TimeSteppingSimulator, scenes
    assert torch.cuda.is_available(), "bench.py needs a CUDA device: the hot path has no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    ne = args.envs_per_gpu
    scene = scenes.small_lcp_batch(ne, seed=0xB200 + rank)      # every rank owns its own envs (contiguous shard of the job)
    if args.min_step == "default":
        scene.min_step_size_env = None
    sim = TimeSteppingSimulator(scene, device=local_rank)
    stream = torch.cuda.current_stream()
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sim.step(DT, args.preroll)
    for _ in range(max(args.warmup, 3)):
        flush.fill_(1)
        sim.step(DT, 1)
    barrier()
    sim.reset_counters()
    launches0 = sim.launch_count()
    q0, v0 = sim.get_state()            # the state the timed region starts from (also feeds the CPU baseline)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    sampler = ClockSampler(local_rank)
    sampler.start()
    barrier()
    wall0 = time.perf_counter()
    for a, b in ev:
        flush.fill_(0)                   # L2 flush, outside the timed events
        a.record(stream)
        sim.step(DT, 1)                  # the hot path: advance / impact-class / finish kernels of one step
        b.record(stream)
    barrier()
    wall = time.perf_counter() - wall0
    clocks = sampler.stop()
    kernel_ms = [a.elapsed_time(b) for a, b in ev]
    t_dev = sum(kernel_ms) * 1e-3
    cnt = sim.counters()
    r_cnt = dict(cnt)
    launches = sim.launch_count() - launches0
    # ---- end to end through the public API with HOST buffers: H2D state, step, D2H state, every step ----
    qh = torch.from_numpy(q0).pin_memory()
    vh = torch.from_numpy(v0).pin_memory()
    qd, vd = torch.empty_like(qh, device=dev), torch.empty_like(vh, device=dev)
    qo, vo = torch.empty_like(qh).pin_memory(), torch.empty_like(vh).pin_memory()
    h2d = qh.numel() * 8 + vh.numel() * 8
    for _ in range(2):
        qd.copy_(qh, non_blocking=True); vd.copy_(vh, non_blocking=True)
        sim.set_state_dev(qd, vd); sim.step(DT, 1); sim.get_state_dev(qd, vd)
        qo.copy_(qd, non_blocking=True); vo.copy_(vd, non_blocking=True)
    barrier()
    e0 = time.perf_counter()
    for _ in range(args.steps):
        qd.copy_(qh, non_blocking=True); vd.copy_(vh, non_blocking=True)
        sim.set_state_dev(qd, vd)
        sim.step(DT, 1)
        sim.get_state_dev(qd, vd)
        qo.copy_(qd, non_blocking=True); vo.copy_(vd, non_blocking=True)
        torch.cuda.synchronize()
        qh, qo = qo, qh                  # next step starts from this step's result
        vh, vo = vo, vh
    barrier()
    t_e2e = time.perf_counter() - e0
    # ---- max over ranks, sums of counters ----
    tt = torch.tensor([t_dev, t_e2e, wall], dtype=torch.float64, device=dev)
    cc = torch.tensor([cnt[k] for k in ("env_steps", "mini_steps", "lcp_solves", "pivots", "pivot_flops", "lcp_failures",
                                        "lcp_fast_calls", "lemke_calls", "contacts")], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dist.all_reduce(cc, op=dist.ReduceOp.SUM)
    t_dev, t_e2e, wall = tt.tolist()
    env_steps, mini_steps, lcp_solves, pivots, pivot_flops, failures, fast_calls, lemke_calls, contacts = cc.tolist()
    if rank == 0:
        total_envs = ne * world
        hbm_peak, hbm_src, fp64_peak, fp64_src = _peaks()
        k_ms = float(np.mean(kernel_ms))
        # roofline of the dominant (only) kernel, per launch on this rank
        alg_bytes = BYTES_PER_ENV_STEP * ne
        alg_flops = (r_cnt["pivot_flops"] + r_cnt["assembly_flops"]) / args.steps
        achieved_gbs = alg_bytes / (k_ms * 1e-3) / 1e9
        achieved_tf = alg_flops / (k_ms * 1e-3) / 1e12
        out = {
            "metric": "env_steps_per_s", "value": total_envs * args.steps / t_dev, "unit": "env-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": 1e3 * t_dev / args.steps,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "envs_per_gpu": ne, "dt": DT, "preroll_steps": args.preroll, "seed": "0xB200+rank",
                       "min_step_size": "boxes 1e-3 (test/box.xml), balls sqrt(eps) (bouncing-ball.xml)" if args.min_step == "scene" else "sqrt(eps) everywhere",
                       "impact_model": "QP-as-LCP (default build)", "stabilization": "off (max-iterations=0)",
                       "l2": "flushed between timed steps (256 MiB write outside the events)", "parallelism": f"envs sharded x{world}"},
            "lcp_solves_per_s": lcp_solves / t_dev,
            "mini_steps_per_step": mini_steps / max(env_steps, 1.0), "lcp_solves_per_env_step": lcp_solves / max(env_steps, 1.0),
            "pivots_per_solve": pivots / max(lcp_solves, 1.0), "lemke_calls": lemke_calls, "lcp_fast_calls": fast_calls,
            "lcp_failures": failures, "contacts_per_env_step": contacts / max(env_steps, 1.0), "ca_iterations_per_env_step": r_cnt["ca_iterations"] / max(r_cnt["env_steps"], 1),
            "wall_s_timed_region": wall,
            "e2e": {"value": total_envs * args.steps / t_e2e, "unit": "env-steps/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": h2d,
                    "how": "pinned host q,v -> device -> b200moby_set_state_dev -> step -> get_state_dev -> pinned host, every step"},
            "gpu_launches": launches * world,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s", "frac": achieved_gbs / hbm_peak,
                         "traffic": None, "kernel": "step_warp_kernel", "kernel_ms": k_ms, "peak_source": hbm_src,
                         "note": "the fused env-step is not HBM-bound (208 algorithmic B per env-step); see fp64",
                         "fp64": {"achieved": achieved_tf, "peak": fp64_peak, "unit": "TFLOP/s", "frac": achieved_tf / fp64_peak,
                                  "peak_source": fp64_src,
                                  "flops": "SURVEY 8(d): sum pivots*2n(n+1) + F_delassus + F_apply per solve + F_fd + F_narrow per mini-step, all from recorded counts"}},
        }
        if not args.no_cpu_baseline:
            n_cpu, s_cpu = 512, 40
            v1, l1, el1, _ = _cpu_baseline(scene, q0, v0, n_cpu, s_cpu, 1)
            cores = os.cpu_count() or 1
            vn, ln, eln, _ = _cpu_baseline(scene, q0, v0, n_cpu * 4, s_cpu, cores)
            out["cpu_baseline"] = {"value": v1, "unit": "env-steps/s", "cores": 1, "kind": "port",
                                   "sample": f"first {n_cpu} envs of rank 0's batch from the same pre-rolled state, {s_cpu} steps, "
                                             f"1 thread ({el1:.1f} s); oracle/ restatement (the reference cannot be built here)",
                                   "lcp_solves_per_s": l1,
                                   "all_cores": {"value": vn, "cores": cores, "sample": f"first {n_cpu * 4} envs, {s_cpu} steps ({eln:.1f} s)"}}
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
