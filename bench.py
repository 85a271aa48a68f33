#!/usr/bin/env python
"""Benchmark of the hot path: converged trajectories / second of the batched trust-region SQP
(BASELINE.json metric), one process per GPU.

  python bench.py --gpus N --steps K --warmup W        # CUDA path (this repo)
  python bench.py --impl reference ...                  # CPU baseline: the oracle restatement of the
                                                        # reference's CPU path on all host threads

A "step" = one complete tb200_solve_batch over one batch of synthetic problems: 1024 x 7-DOF x 30 waypoints
with the 8-sphere discrete collision constraint (BASELINE.json configs[2], the "collision-constrained"
workload the metric's target is stated on; --config cfg1 drops the collision term = configs[1]).
`value` is measured with the inputs already resident in HBM (tb200_solve_batch_resident); `e2e` is the
same metric through the public API with HOST buffers (H2D of the per-trajectory inputs and D2H of the
results inside the timed region).  Every step uses a different synthetic batch (fresh seeds), so nothing
is cached between timed iterations; the working set of one step (~0.6 GB of convexification rows + QP
workspace at B=1024 with collision) is larger than L2 (126 MB).
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
sys.path.insert(0, ROOT)

from trajopt_b200 import problems  # noqa: E402

METRIC = "converged trajectories/sec (7-DOF x 30 wp, batch 1024)"
# dram__bytes_read.sum + dram__bytes_write.sum of one full-batch convexify launch (ncu --set full, profiles/)
TRAFFIC = {"cfg2": 103.0e6}  # 4.2 MB read + 98.8 MB written (r01; the 126 MB L2 still holds part of the rows at kernel end)


def make_batch(config, batch, seed):
    if config == "cfg1":
        return problems.config1(B=batch, T=30, seed=seed)
    if config == "cfg2":
        return problems.config2(B=batch, T=30, seed=seed)
    if config == "cfg3":  # configs[3] terms at 30 waypoints (the QP kernel holds <= 30 waypoints of 7 joints so far)
        return problems.config3(B=batch, T=30, seed=seed)
    raise SystemExit(f"unknown config {config}")


def workload_name(config, batch):
    if config == "cfg3":
        return (f"batch {batch} x 7-DOF x 30 waypoints, JointVel/JointAcc + CartPose via/terminal constraints + CartVel + "
                "LVS continuous collision (8 sphere obstacles) = configs[3] terms at 30 waypoints")
    extra = " + discrete collision (8 sphere obstacles), safety_margin 0.02" if config == "cfg2" else ""
    return f"batch {batch} x 7-DOF x 30 waypoints, JointVel/JointAcc + CartPose terminal constraint{extra}"


def host_threads():
    """Host threads of the CPU legs: one per physical core this process may run on (torchrun exports OMP_NUM_THREADS=1,
    which must not shrink the CPU baseline).  Measured on the GPU box (2 x 64 hardware threads), full batch of 1024:
    64 threads 6.3 s, 128 threads 8.5 s - the oracle is bound by its allocator and caches, SMT siblings only hurt."""
    if os.environ.get("TB200_CPU_THREADS"):
        return int(os.environ["TB200_CPU_THREADS"])
    try:
        logical = len(os.sched_getaffinity(0))
    except AttributeError:
        logical = os.cpu_count() or 1
    try:
        import psutil
        physical = psutil.cpu_count(logical=False) or logical
    except ImportError:
        physical = logical
    return max(1, min(logical, physical))


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, index):
        self.index = index
        self.samples = []
        self.proc = None

    def start(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-i",
                                          str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons}


def run_reference(args, rank, world):
    """CPU arm: the reference's own CPU path restated (oracle/), OpenMP over trajectories on all host threads.
    Each step is a bounded sample of the same workload (the reference cannot be compiled here: no Eigen /
    OSQP / tesseract)."""
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle_lib
    oracle_lib.build()
    threads = host_threads()
    sample = args.cpu_sample
    times, conv = [], []
    for it in range(args.warmup + args.steps):
        desc = make_batch(args.config, sample, problems.SEED + 1000 + it)
        t0 = time.perf_counter()
        r = oracle_lib.solve_batch(desc, n_threads=threads)
        dt = time.perf_counter() - t0
        if it >= args.warmup:
            times.append(dt)
            conv.append(int((r["status"] == 0).sum()))
    value = sum(conv) / sum(times)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * float(np.mean(times)), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.batch), "timing": "fresh synthetic batch every step"},
            "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": threads, "kind": "port",
                             "sample": f"{sample} trajectories of the same workload per step, OpenMP over trajectories"},
            "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="cuda", choices=["cuda", "reference"])
    ap.add_argument("--config", default="cfg2", choices=["cfg1", "cfg2", "cfg3"])
    ap.add_argument("--batch", type=int, default=1024, help="trajectories per GPU (weak scaling)")
    ap.add_argument("--cpu-sample", type=int, default=1024,
                    help="trajectories per CPU baseline step (default: one whole batch, so that the CPU path is bound by "
                         "its longest trajectory exactly as the GPU path is)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl cuda needs a CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    from trajopt_b200 import api, capi

    # each rank owns `batch` independent trajectories (shards never interact: no data-path collective)
    total = args.warmup + args.steps
    batches = [make_batch(args.config, args.batch, problems.SEED + 1 + 7919 * (it * world + rank)) for it in range(total)]
    prob = api.Problem(batches[0], device=local_rank)
    pinned = [dict(init=torch.from_numpy(b.init_traj).pin_memory(), tgt=torch.from_numpy(b.cart_targets).pin_memory(),
                   obs=None if b.obstacles is None else torch.from_numpy(b.obstacles).pin_memory()) for b in batches]

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def step(it, resident):
        pb = pinned[it]
        ptr = lambda t: None if t is None else capi.C.cast(t.data_ptr(), capi.C.POINTER(capi.C.c_double))
        if resident:
            prob._check(prob.lib.tb200_problem_set_inputs(prob.handle, ptr(pb["init"]), ptr(pb["tgt"]), ptr(pb["obs"])))
            sync()
            t0 = time.perf_counter()
            prob.solve_resident()
            sync()
            dt = time.perf_counter() - t0
            tm = prob.timing()
            res = prob.fetch()
        else:
            sync()
            t0 = time.perf_counter()
            prob._check(prob.lib.tb200_problem_set_inputs(prob.handle, ptr(pb["init"]), ptr(pb["tgt"]), ptr(pb["obs"])))
            res = prob.solve()
            sync()
            dt = time.perf_counter() - t0
            tm = res["timing"]
        return dt, tm, res

    sampler = ClockSampler(local_rank)
    # ---- resident leg (value) -----------------------------------------------------------------------------
    for it in range(args.warmup):
        step(it, True)
    if rank == 0:
        sampler.start()
    dts, dev_ms, conv, tms, ktm = [], [], [], [], []
    for it in range(args.warmup, total):
        dt, tm, res = step(it, True)
        dts.append(dt)
        dev_ms.append(tm["total_ms"])
        conv.append(int((res["status"] == 0).sum()))
        tms.append(tm)
        # the convexify kernel alone, every trajectory active, at this step's solution (new data every launch;
        # one launch writes ~135 MB > L2): the launch the roofline below is quoted on
        ktm.append(prob.convexify_timed(res["x"]))
    clocks = sampler.stop() if rank == 0 else None
    # ---- end-to-end leg (host buffers, H2D + D2H inside the timed region) ------------------------------------
    e2e_dts, e2e_conv, h2d, d2h = [], [], 0, 0
    for it in range(args.warmup, total):
        dt, tm, res = step(it, False)
        e2e_dts.append(dt)
        e2e_conv.append(int((res["status"] == 0).sum()))
        h2d, d2h = tm["h2d_bytes"], tm["d2h_bytes"]

    def reduce_max(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)  # the one collective of the path: converged-trajectory count
        return float(t.item())

    dev_total_s = reduce_max(sum(dev_ms) / 1e3)     # device time (CUDA events on the solver stream), max over ranks
    wall_total_s = reduce_max(sum(dts))
    conv_total = reduce_sum(float(sum(conv)))
    e2e_total_s = reduce_max(sum(e2e_dts))
    e2e_conv_total = reduce_sum(float(sum(e2e_conv)))
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    value = conv_total / dev_total_s
    # ---- roofline of the convexify kernel (HBM bound; algorithmic bytes per launch: DESIGN.md §4) -------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst)"
    conv_ms = sum(t["convexify_ms"] for t in tms)
    conv_launches = sum(t["convexify_launches"] for t in tms)
    conv_bytes = sum(t["convexify_bytes"] for t in tms)  # algorithmic bytes of the trajectories actually convexified
    k_ms = sum(t["convexify_ms"] for t in ktm)
    k_bytes = sum(t["convexify_bytes"] for t in ktm)
    achieved = k_bytes / (k_ms * 1e-3) / 1e9 if k_ms > 0 else 0.0
    roofline = {"bound": "hbm", "kernel": "eval_convexify_decide_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": TRAFFIC.get(args.config), "peak_source": peak_src,
                "scope": "one full-batch launch per timed step (all trajectories active), CUDA events on the launching stream",
                "avg_launch_us": 1e3 * k_ms / max(len(ktm), 1), "algorithmic_bytes_per_launch": k_bytes / max(len(ktm), 1),
                "in_step": {"share_of_step": conv_ms / (sum(dev_ms)), "evaluations": conv_launches,
                            "note": "inside a solve the same code runs as a step of the persistent solve_kernel, one "
                                    "trajectory per CTA at a time (share = SM time in evaluation steps, %globaltimer)"}}
    qp_ms = sum(t["qp_ms"] for t in tms)
    line = {"metric": METRIC, "value": value, "unit": "trajectories/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": 1e3 * dev_total_s / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": workload_name(args.config, args.batch), "global_batch": args.batch * world,
                       "parallelism": f"batch sharded over {world} GPU(s), no data-path collective",
                       "timing": "fresh synthetic batch every step; per-step working set > L2"},
            "converged_fraction": conv_total / (args.batch * world * args.steps),
            "wall_ms_per_step": 1e3 * wall_total_s / args.steps,
            "e2e": {"value": e2e_conv_total / e2e_total_s, "unit": "trajectories/s", "h2d_bytes_per_step": int(h2d),
                    "d2h_bytes_per_step": int(d2h)},
            # per solve: reset_state_kernel, eval_convexify_decide_kernel (initial evaluation), solve_kernel (persistent);
            # plus the stand-alone convexify launch the roofline is quoted on
            "gpu_launches": 4 * len(tms),
            "roofline": roofline,
            "qp_steps": {"share_of_step": qp_ms / sum(dev_ms), "qp_solves": int(sum(t["qp_launches"] for t in tms)),
                         "note": "QP steps of solve_kernel (ADMM): shared-memory/latency bound, see profiles/ for achieved occupancy"},
            "clocks": clocks}
    if not args.no_cpu_baseline:
        sys.path.insert(0, os.path.join(ROOT, "tests"))
        import oracle_lib  # CPU baseline leg: the oracle is the timed CPU path, never part of the product
        threads = host_threads()
        desc = make_batch(args.config, args.cpu_sample, problems.SEED + 999)
        t0 = time.perf_counter()
        r = oracle_lib.solve_batch(desc, n_threads=threads)
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": float((r["status"] == 0).sum() / dt), "unit": "trajectories/s", "cores": threads,
                                "kind": "port", "sample": f"{args.cpu_sample} trajectories of the same workload, {dt:.1f} s"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
