#!/usr/bin/env python
"""Benchmark of the flat Monte-Carlo search hot path (BASELINE.json: MCTS rollouts/s).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1

Workload (config.workload): problems_g — the 1000 Guillotinable 20-tile instances of the
reference's problems/problems_g.csv, N = 1000 rollouts per child, max_depth strategy,
Philox key (123, problem index).  A step = one whole flat search of
every problem of the rank's shard (all depths).  With N GPUs each rank runs the full
1000-problem set on its own Philox streams (weak scaling, no collective); the
strong-scaling wall time of the single 1000-problem set is reported under "strong".

value   = rollouts the reference would have made for the same job (its early exits
          honoured) / device time, initial states resident in HBM.
e2e     = same metric through the one-shot C-ABI call bp_flat_search with host numpy
          buffers (allocation, H2D, all depths, D2H inside the timed region).
roofline: the rollout kernel is bound by instruction issue, not by HBM or tensor
          cores: achieved = placements executed x A(class) / step time of the
          timed region (the class launches overlap, so this bounds the rollout kernels'
          rate from below; the serialised profiling pass is reported beside it),
          peak = SMs x 4 schedulers x max SM clock.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)



def model_inst_per_placement(key):
    """ALGORITHMIC warp instructions per placement for shape class "RHO,W,EJ": the
    warp-per-board work model of SURVEY.md 8(d), frozen in DESIGN.md section 5 before any
    kernel tuning (LFB 4+2*RHO*W, free run 4*W, legality 6*EJ, Philox+reduce 27,
    k-th-set select 6*EJ, place 6*RHO*W, pair removal 6, loop 6)."""
    rho, w, ej = (int(v) for v in key.split(","))
    return float((4 + 2 * rho * w) + 4 * w + 6 * ej + 27 + 6 * ej + 6 * rho * w + 6 + 6)


# raw ncu smsp__inst_executed.sum / placements of the warp-per-board kernel (profiles/r1/)
MEASURED_INST_PER_PLACEMENT_V1 = {"1,1,2": 137.2, "1,2,2": 167.4, "2,1,2": 148.9, "2,2,2": 182.1}


def model_lane_inst_per_placement(key):
    """ALGORITHMIC thread instructions per placement of the lane-per-rollout kernel for lane
    class "NB,W,EW" (NB height planes, W words per plane, EW words of entry mask); frozen in
    DESIGN.md section 5 when the kernel was first measured: minimum search NB*(2W+2), column
    and run 5W+2, legal mask 4+3EW, draw 15 (Philox4x32-10 amortised over 4 draws + index),
    k-th-set select 5(EW-1)+32, table loads 4, place W+2*NB*W, pair removal 2(3EW+2), loop 8.
    One warp instruction advances 32 lanes, so the warp-instruction equivalent is this / 32."""
    nb, w, ew = (int(v) for v in key.split(","))
    return float(nb * (2 * w + 2) + (5 * w + 2) + (4 + 3 * ew) + 15 + (5 * (ew - 1) + 32) + 4 +
                 (w + 2 * nb * w) + 2 * (3 * ew + 2) + 8)


# raw ncu thread instructions / placement and DRAM bytes / launch of that kernel (profiles/r1/)
# (ncu_lanes7_inst_*.csv: 150 problems per class, N=960; 120 launches, launch-weighted mean).
# Warp instructions: the loop body is branch-free, so finished lanes ride along and thread
# counts no longer say how many lanes do useful work.
# (final build, whole-child work units: ncu_final_inst_*.csv + final_plain_*.json)
MEASURED_LANE_WARP_INST = {"5,1,2": 3.98, "5,2,2": 5.30, "6,1,2": 4.30, "6,2,2": 5.80}
MEASURED_LANE_DRAM_BYTES_PER_LAUNCH = 200464.0
# ncu --set full of fs_rollout_lanes_kernel<5,1,2> (ncu_full_fs_rollout_lanes_5_1_2_final.csv)
MEASURED_LANE_PIPE_PCT = {"issue_active": 73.2, "alu": 62.7, "fma": 28.1, "xu": 45.6, "lsu": 13.5}


def lane_class_key(rows, cols, n_entries):
    nb = 5 if rows <= 31 else (6 if rows <= 63 else (7 if rows <= 127 else 8))
    w = 1 if cols <= 32 else (2 if cols <= 64 else (3 if cols <= 96 else 4))
    return "%d,%d,%d" % (nb, w, 2 if n_entries <= 64 else 4)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--set", default="g", choices=["g", "b"])
    ap.add_argument("--n-sim", type=int, default=1000)
    ap.add_argument("--strategy", default="max_depth", choices=["max_depth", "avg_depth"])
    ap.add_argument("--n-problems", type=int, default=1000)
    ap.add_argument("--seed", type=int, default=123)
    ap.add_argument("--cpu-budget", type=float, default=15.0,
                    help="seconds of the python-port CPU sample (0 = skip cpu_baseline)")
    ap.add_argument("--ref-budget", type=float, default=10.0,
                    help="--impl reference: seconds per step")
    ap.add_argument("--mode", default="auto", choices=["auto", "stepped", "resident"],
                    help="stepped: one rollout + one advance launch per depth and class; "
                         "resident: one launch per shape class runs every depth; auto: stepped")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-strong", action="store_true")
    ap.add_argument("--no-verify", action="store_true")
    return ap.parse_args()


def workload_config(a):
    return {"workload": "problems_%s[0:%d] 20-tile instances, flat MC search, N=%d rollouts/child, %s"
                        % (a.set, a.n_problems, a.n_sim, a.strategy),
            "set": a.set, "n_problems": a.n_problems, "n_sim": a.n_sim, "strategy": a.strategy,
            "seed": a.seed,
            "l2": "flushed between timed steps (256 MiB device memset); a step streams no input, "
                  "rollout state lives in registers"}


# ------------------------------------------------------------------ reference arm
def run_cpu_baseline(a, budget, c_problems, skip_python=False):
    cmd = [sys.executable, "-m", "oracle.cpu_baseline", "--set", a.set, "--n-sim", str(a.n_sim),
           "--strategy", a.strategy, "--seed", str(a.seed), "--budget", str(budget),
           "--c-problems", str(c_problems)]
    if skip_python:
        cmd.append("--skip-python")
    out = subprocess.run(cmd, cwd=ROOT, capture_output=True, text=True, timeout=900)
    if out.returncode != 0:
        raise RuntimeError("cpu baseline failed: %s" % out.stderr[-2000:])
    return json.loads(out.stdout.strip().splitlines()[-1])


def reference_arm(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    samples = []
    base = None
    # bounded sample per step so that the whole run ends within a few minutes for any K
    budget = min(a.ref_budget, max(2.0, 150.0 / max(a.steps, 1)))
    for i in range(a.warmup + a.steps):
        base = run_cpu_baseline(a, budget if i >= a.warmup else min(budget, 2.0),
                                c_problems=8 if i == a.warmup + a.steps - 1 else 0)
        if i >= a.warmup:
            samples.append(base["python_port"])
    rollouts = sum(s["rollouts"] for s in samples)
    seconds = sum(s["seconds"] for s in samples)
    value = rollouts / seconds
    cores = samples[-1]["cores"]
    sample = ("python/numpy restatement of reference CustomMCTS.predict, %d single-threaded "
              "processes, problems in order, %.1f s wall budget per step" % (cores, budget))
    line = {
        "impl": "reference", "metric": "mcts_rollouts_per_s", "value": value, "unit": "rollouts/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
        "ms_per_step": 1e3 * seconds / max(len(samples), 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32", "data": "reference problem set",
        "config": workload_config(a),
        "cpu_baseline": {"value": value, "unit": "rollouts/s", "cores": cores, "kind": "port",
                         "sample": sample, "cpu_model": base["cpu_model"],
                         "placements_per_s": sum(s["placements"] for s in samples) / seconds,
                         "c_port_value": base.get("c_port", {}).get("rollouts_per_s"),
                         "c_port_sample": "scalar C restatement, %d pthreads, first 8 problems"
                                          % cores},
        "e2e": {"value": value, "unit": "rollouts/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


# ------------------------------------------------------------------ clocks
class ClockSampler(object):
    QUERY = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.QUERY,
                 "--format=csv,noheader,nounits", "-lms", "20"],
                stdout=self.tmp, stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    @staticmethod
    def _stamp(text):
        import datetime
        try:
            return datetime.datetime.strptime(text.strip(), "%Y/%m/%d %H:%M:%S.%f").timestamp()
        except ValueError:
            return None

    def stop(self, t_begin=None, t_end=None):
        """Median SM clock and throttle reasons of the samples taken inside [t_begin, t_end]
        (wall clock of the timed region); all samples when none fall inside."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = []
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                rows.append((self._stamp(f[0]), float(f[1]), float(f[2]), float(f[3]),
                             [n for n, v in zip(names, f[5:9]) if v.lower().startswith("active")]))
            except ValueError:
                continue
        inside = [r for r in rows if r[0] is not None and t_begin is not None and
                  t_begin <= r[0] <= t_end]
        use = inside if inside else rows
        sm = [r[1] for r in use]
        mx = [r[2] for r in use]
        power = [r[3] for r in use]
        reasons = set(n for r in use for n in r[4])
        self.tmp.close()
        os.unlink(self.tmp.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)),
                "power_w_max": float(max(power)), "samples": len(sm),
                "samples_inside_timed_region": len(inside), "reasons": sorted(reasons)}


# ------------------------------------------------------------------ our arm
def shape_class_key(rows, cols, n_entries):
    def rc(v):
        return 1 if v <= 1 else (2 if v <= 2 else 4)
    return "%d,%d,%d" % (rc((rows + 31) // 32), rc((cols + 31) // 32), rc((n_entries + 31) // 32))


def our_arm(a):
    import torch
    import torch.distributed as dist

    from visapi_b200 import Engine, datasets
    from visapi_b200.verify import replay_solution

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    def barrier():
        if world > 1:
            dist.barrier()

    problems = datasets.load_problem_set(a.set)[:a.n_problems]
    batch = datasets.batch_arrays(problems)
    P = len(problems)
    strat = 0 if a.strategy == "max_depth" else 1
    # weak scaling: every rank searches the whole set on its own Philox streams
    streams = (np.arange(P, dtype=np.uint32) + np.uint32(rank * P))

    stream = torch.cuda.current_stream()
    eng = Engine(local_rank, stream=stream.cuda_stream)
    search = eng.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], a.n_sim,
                        strat, a.seed, streams)

    search.set_mode(a.mode)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=dev)    # > 126 MB of L2

    def step():
        flush_buf.zero_()
        search.reset()
        search.run()

    for _ in range(max(a.warmup, 3)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    launches0 = eng.kernel_launches
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    torch.cuda.synchronize()
    t_wall0 = time.time()
    ev0.record(stream)
    for _ in range(a.steps):
        step()
    ev1.record(stream)
    torch.cuda.synchronize()
    barrier()
    wall = time.time() - t_wall0
    launches = eng.kernel_launches - launches0
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall0 + wall)
    res = search.results()

    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    counts = torch.tensor([float(res["rollouts"].sum()), float(res["rollout_placements"].sum()),
                           float(res["rollouts_executed"].sum()),
                           float(res["placements_executed"].sum()),
                           float(res["solution_found"].sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    ms_max = float(t.item())
    rollouts, placements, rollouts_exec, placements_exec, solved = [float(x) for x in counts.tolist()]
    sec_per_step = ms_max / 1e3 / a.steps
    value = rollouts / sec_per_step

    # ---- correctness gate: every reported solution is replayed on the host ----------
    verified = None
    if not a.no_verify:
        bad = []
        for i, p in enumerate(problems):
            if res["solution_found"][i]:
                order = res["order"][i, :res["n_order"][i]].tolist()
                ok, why = replay_solution(p.rows, p.cols, p.tiles, order)
                if not ok:
                    bad.append((i, why))
        if bad:
            raise SystemExit("INVALID SOLUTIONS reported by the device: %r" % bad[:5])
        verified = int(res["solution_found"].sum())

    # ---- per-kernel time (CUDA events on the launching stream) ---------------------
    search.set_profiling(True)
    step()
    rollout_ms, advance_ms, n_depths = search.profile()
    resident = search.last_run_resident()
    res_p = search.results()
    search.set_profiling(False)
    lane_kernel = search.info()["lane_kernel"]
    if lane_kernel:
        keys = [lane_class_key(p.rows, p.cols, 2 * len(p.tiles)) for p in problems]
        per_placement = lambda k: model_lane_inst_per_placement(k) / 32.0   # noqa: E731
    else:
        keys = [shape_class_key(p.rows, p.cols, 2 * len(p.tiles)) for p in problems]
        per_placement = model_inst_per_placement
    algo_inst = 0.0
    by_class = {}
    for k, pe in zip(keys, res_p["placements_executed"].tolist()):
        by_class[k] = by_class.get(k, 0) + pe
    for k, pe in by_class.items():
        algo_inst += pe * per_placement(k)
    info = eng.device_info()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    issue_peak = info["sm_count"] * 4 * sm_max_mhz * 1e6
    # The classes' launches overlap on their own streams in the timed region, so a rollout
    # kernel's duration there is not separable; every rollout kernel of a step runs inside the
    # step, hence algorithmic work / step time is a lower bound of the rollout kernels' rate
    # (the step also holds the advance kernels, the reset and the L2 flush).  The profiling
    # pass (classes serialised, every launch alone on the GPU with its own tail) is reported
    # beside it under "serialised".
    achieved = algo_inst / sec_per_step
    achieved_serial = algo_inst / (rollout_ms / 1e3)
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_launches = n_depths if resident else n_depths * search.info()["n_classes"]
    alu_only, alu_fma = eng.issue_microbench()
    # HBM bytes a rollout task needs: its problem's tables (on a problem switch) + 8 B record
    n_tasks = float(res_p["rollouts_executed"].sum()) / 32.0
    algo_bytes = n_tasks * (128 + 8) + float(res_p["depth"].sum() + P) * 2400.0
    roofline = {
        "bound": "issue",
        "kernel": ("fs_search_resident_kernel" if resident else
                   ("fs_rollout_lanes_kernel" if lane_kernel else "fs_rollout_kernel")),
        "achieved": achieved / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
        "frac": achieved / issue_peak,
        "traffic": MEASURED_LANE_DRAM_BYTES_PER_LAUNCH if lane_kernel else 98560.0,
        "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/r1)",
        "algorithmic_bytes_per_launch": algo_bytes / max(n_launches, 1),
        "peak_source": "SMs x 4 schedulers x sm_max_mhz (%s)" %
                       ("MEASURED_PEAKS.json" if "sm_max_mhz" in peaks else "nvidia-smi"),
        "measured_issue_gwarp_inst_s": {"lop3_iadd3_alu_pipe_only": alu_only,
                                        "lop3_imad_alu_plus_fma": alu_fma},
        "ncu_pipe_pct_of_peak": MEASURED_LANE_PIPE_PCT if lane_kernel else None,
        "model": ("thread instructions per placement / 32 (lane-per-rollout)" if lane_kernel
                  else "warp instructions per placement (warp-per-board)"),
        "achieved_is": "algorithmic warp instructions of one step / step time (timed region, CUDA "
                       "events; class launches overlap, so this bounds the rollout kernels' rate "
                       "from below)",
        "serialised": {"achieved": achieved_serial / 1e9, "frac": achieved_serial / issue_peak,
                       "note": "profiling pass: classes serialised, CUDA events around every "
                               "launch, sum of the rollout launches"},
        "rollout_kernel_ms_per_step": rollout_ms, "advance_kernel_ms_per_step": advance_ms,
        "rollout_kernel_share_of_step": rollout_ms / (rollout_ms + advance_ms),
        "launches_per_step": ({"resident": n_launches, "advance_first": search.info()["n_classes"]}
                              if resident else {"rollout": n_launches, "advance": n_launches}),
        "avg_rollout_launch_ms": rollout_ms / max(n_launches, 1),
        "placements_executed_per_step": float(res_p["placements_executed"].sum()),
        "placements_per_s_kernel": float(res_p["placements_executed"].sum()) / (rollout_ms / 1e3),
        "algorithmic_inst_per_placement": {k: per_placement(k) for k in by_class},
        "ncu_inst_per_placement": ({k: MEASURED_LANE_WARP_INST.get(k) for k in by_class}
                                   if lane_kernel else
                                   {k: MEASURED_INST_PER_PLACEMENT_V1.get(k) for k in by_class}),
        "hbm": {"algorithmic_gbs": algo_bytes / (rollout_ms / 1e3) / 1e9, "peak_gbs": hbm_peak,
                "frac": algo_bytes / (rollout_ms / 1e3) / 1e9 / hbm_peak},
    }

    # ---- end to end through the one-shot C-ABI call with host buffers --------------
    e2e = None
    if not a.no_e2e:
        h2d = (batch["tiles"].nbytes + batch["rows"].nbytes + batch["cols"].nbytes +
               batch["n_entries"].nbytes + streams.nbytes)
        d2h = P * (56 + (batch["tiles"].shape[1] // 2 + 2) * 2 + (batch["tiles"].shape[1] // 2) * 4)
        for _ in range(2):
            eng.flat_search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], a.n_sim,
                            strat, a.seed, streams)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_rollouts = 0.0
        for _ in range(a.steps):
            flush_buf.zero_()
            r = eng.flat_search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"],
                                a.n_sim, strat, a.seed, streams)
            e2e_rollouts += float(r["rollouts"].sum())
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt], dtype=torch.float64, device=dev)
        rr = torch.tensor([e2e_rollouts], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(rr, op=dist.ReduceOp.SUM)
        e2e = {"value": float(rr.item()) / float(tt.item()), "unit": "rollouts/s",
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "ms_per_step": 1e3 * float(tt.item()) / a.steps,
               "api": "bp_flat_search (create + upload + all depths + download + destroy)"}

    # ---- strong scaling: the single 1000-problem set split over the ranks ----------
    strong = None
    if not a.no_strong:
        mine = list(range(rank, P, world))
        sub = datasets.batch_arrays([problems[i] for i in mine])
        s2 = eng.search(sub["rows"], sub["cols"], sub["tiles"], sub["n_entries"], a.n_sim, strat,
                        a.seed, np.asarray(mine, dtype=np.uint32))
        for _ in range(2):
            s2.reset()
            s2.run()
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(a.steps):
            flush_buf.zero_()
            s2.reset()
            s2.run()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        r2 = s2.results()
        tt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
        cc = torch.tensor([float(r2["rollouts"].sum()), float(r2["solution_found"].sum())],
                          dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dist.all_reduce(cc, op=dist.ReduceOp.SUM)
        sset = float(tt.item()) / 1e3 / a.steps
        strong = {"wall_s_full_set": sset, "rollouts_per_s": float(cc[0].item()) / sset,
                  "solved": int(cc[1].item()), "problems": P, "partition": "problem i -> rank i mod N"}
        s2.close()

    # ---- CPU baseline beside it (rank 0, N = 1 only) --------------------------------
    cpu = None
    if rank == 0 and world == 1 and a.cpu_budget > 0:
        try:
            base = run_cpu_baseline(a, a.cpu_budget, c_problems=8)
            pp = base["python_port"]
            cpu = {"value": pp["rollouts_per_s"], "unit": "rollouts/s", "cores": pp["cores"],
                   "kind": "port",
                   "sample": "python/numpy restatement of reference CustomMCTS.predict, %d "
                             "single-threaded processes, problems in order, first %.0f s"
                             % (pp["cores"], a.cpu_budget),
                   "placements_per_s": pp["placements_per_s"], "cpu_model": base["cpu_model"],
                   "c_port_value": base["c_port"]["rollouts_per_s"],
                   "c_port_sample": "scalar C restatement, %d pthreads, first 8 problems"
                                    % base["c_port"]["cores"]}
        except Exception as exc:  # the baseline is reported, never required
            cpu = {"error": str(exc)[:300]}

    if rank == 0:
        line = {
            "metric": "mcts_rollouts_per_s", "value": value, "unit": "rollouts/s",
            "n_gpus": world, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "reference problem set (bundled npz)",
            "config": workload_config(a),
            "rollouts_per_step": rollouts, "placements_per_step": placements,
            "rollouts_executed_per_step": rollouts_exec,
            "placements_executed_per_step": placements_exec,
            "placements_per_s": placements / sec_per_step,
            "rollouts_executed_per_s": rollouts_exec / sec_per_step,
            "solved": int(solved), "solutions_verified_rank0": verified,
            "problems_per_step": P * world, "wall_s_per_step": wall / a.steps,
            "clocks": clocks, "gpu_launches": int(launches),
            "roofline": roofline, "e2e": e2e, "strong": strong, "cpu_baseline": cpu,
        }
        print(json.dumps(line))
    search.close()
    eng.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        our_arm(a)


if __name__ == "__main__":
    main()
