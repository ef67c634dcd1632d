#!/usr/bin/env python
"""Benchmark of the flat Monte-Carlo search hot path (BASELINE.json: MCTS rollouts/s).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1

Workload (config.workload): problems_g -- the 1000 Guillotinable 20-tile instances of the
reference's problems/problems_g.csv, N = 1000 rollouts per child, max_depth strategy,
Philox key (123, problem index).  A step = one whole flat search (all
depths) of THE 1000-problem set.  With N GPUs the set is sharded, problem i -> rank i mod N, no
collective on the data path (BASELINE config 2): value = the set's rollouts / time of the
slowest rank, "scaling": "strong".  Every rank also searches the whole set once and checks that
its shard's rows are bit-identical to the single-GPU result.  The replica throughput (every rank
searching a full set of its own) is reported under "weak".

value   = rollouts the reference would have made for the same job (its early exits
          honoured) / device time, initial states resident in HBM.
e2e     = same metric through the one-shot C-ABI call bp_flat_search with host numpy
          buffers (allocation, H2D, all depths, D2H inside the timed region).
roofline: the rollout kernel is bound by instruction issue, not by HBM or tensor
          cores: achieved = placements executed x A(class) / step time of the
          timed region, peak = SMs x 4 schedulers x max SM clock; "frac" uses the frozen work
          model A, "frac_min_ops" an implementation-independent minimum operation count.
Sub-records: "root_parallel" (config 4: one 50-tile instance, rollouts of every child split
over the ranks, merged inside the resident kernels over NVLink), "puct_root_parallel" (root
Nsa / Wsa all-reduce), "selfplay" (config 5: lock-step batched self-play).
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


def min_ops_lane_per_placement(key):
    """Implementation-independent lower bound: the fewest 32-bit operations one rollout needs per
    placement on a bit-sliced skyline (NB planes x W words, EW words of entry mask), whatever the
    kernel looks like -- minimum search NB*(W+1) (AND per word, OR-test per plane), first column
    and run 6 + 3(W-1), legal mask 5EW - 1 (two table rows, two ANDs, popcounts, sum), one draw
    at Philox4x32-10 cost 60/4 + 1 (index), k-th legal entry 15 + 2(EW-1) (five halving steps),
    one entry fetch, place 3 + NB*W (footprint, one flip per plane word), pair removal EW,
    bookkeeping 3.  Thread operations; / 32 for the warp-instruction equivalent."""
    nb, w, ew = (int(v) for v in key.split(","))
    return float(nb * (w + 1) + 6 + 3 * (w - 1) + (5 * ew - 1) + 16 + 15 + 2 * (ew - 1) + 1 +
                 (3 + nb * w) + ew + 3)


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
    ap.add_argument("--mode", default="auto", choices=["auto", "stepped", "resident", "hybrid"],
                    help="stepped: one rollout + one advance launch per depth and class; "
                         "resident: one launch per shape class runs every depth; auto: stepped")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-weak", action="store_true", help="skip the replica (weak-scaling) sub-record")
    ap.add_argument("--no-verify", action="store_true")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the root_parallel / puct_root_parallel / selfplay sub-records")
    return ap.parse_args()


def workload_config(a):
    return {"workload": "problems_%s[0:%d] 20-tile instances, flat MC search, N=%d rollouts/child, %s"
                        % (a.set, a.n_problems, a.n_sim, a.strategy),
            "set": a.set, "n_problems": a.n_problems, "n_sim": a.n_sim, "strategy": a.strategy,
            "seed": a.seed, "partition": "problem i -> rank i mod n_gpus (one set, no data-path collective)",
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


def _digest(res, keys=("depth", "solution_found", "n_tiles_placed", "rollouts", "n_order", "order")):
    import hashlib
    h = hashlib.sha256()
    for k in keys:
        h.update(np.ascontiguousarray(res[k]).tobytes())
    return h.hexdigest()[:16]


def _c4_instance(n, rows, cols, seed):
    """50 tiles on rows x cols from the reference's generator."""
    import random
    from visapi_b200 import datasets
    from visapi_b200.data_generator import DataGenerator
    random.seed(seed)
    tiles, board = DataGenerator().gen_tiles_and_board(n, rows, cols, order_tiles=True)
    pool, base = [tuple(int(v) for v in t) for t in tiles], []
    while pool:
        t = pool.pop(0)
        pool.remove((t[1], t[0]))
        base.append(t)
    return datasets.Problem(board.shape[0], board.shape[1], base, "gen%d_%dx%d_s%d" % (n, rows, cols, seed))


def root_parallel_record(eng, torch, dist, world, rank, dev, n_sim=5000, seeds=4, reps=3):
    """Config 4: one hard instance, the rollouts of every child split over the ranks, the per-move
    merge inside the resident kernels over NVLink peer memory.  ms per instance (reset + search,
    CUDA events, max over ranks) and the proof that the result is the single-rank result."""
    from visapi_b200 import datasets, parallel
    out = {"n_sim": n_sim, "tiles": 50, "seeds": seeds, "exchange": "in-kernel peer-memory stores + flags "
           "(CUDA IPC), one resident launch per rank" if world > 1 else "single rank: one resident launch",
           "boards": {}}
    all_same = True
    for (rows, cols) in ((25, 25), (40, 40)):
        times, single_times = [], []
        for seed in range(seeds):
            batch = datasets.batch_arrays([_c4_instance(50, rows, cols, seed)])
            streams = np.array([seed], np.uint32)

            def timed(rp):
                best = None
                for rep in range(reps + 1):                       # the first run warms up
                    torch.cuda.synchronize()
                    if world > 1:
                        dist.barrier()
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rp.run()
                    e1.record()
                    torch.cuda.synchronize()
                    t = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
                    if world > 1:
                        dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    if rep:
                        best = float(t.item()) if best is None else min(best, float(t.item()))
                return best
            rp = parallel.RootParallelSearch(eng, batch, n_sim, 0, 123, streams, exchange="resident")
            times.append(timed(rp))
            digest = _digest(rp.results())
            rp.close()
            # the same instance on ONE rank (every rank does it: also the 1-GPU time beside it)
            single = eng.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], n_sim, 0,
                                123, streams)
            single.set_mode("resident")
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            best = None
            for rep in range(reps + 1):
                single.reset()
                torch.cuda.synchronize()
                e0.record()
                single.run()
                e1.record()
                torch.cuda.synchronize()
                if rep:
                    best = e0.elapsed_time(e1) if best is None else min(best, e0.elapsed_time(e1))
            single_times.append(best)
            all_same = all_same and (_digest(single.results()) == digest)
            single.close()
        out["boards"]["%dx%d" % (rows, cols)] = {
            "ms_per_instance": float(np.mean(times)),
            "ms_per_instance_one_gpu_search_only": float(np.mean(single_times)),
            "speedup_vs_one_gpu": float(np.mean(single_times)) / float(np.mean(times))}
    flag = torch.tensor([1.0 if all_same else 0.0], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    out["digest_equals_single_rank"] = bool(flag.item() == 1.0)
    return out


def puct_root_parallel_record(local_rank, torch, dist, world, rank, dev):
    """every rank searches its own PUCT trees from the same roots, the root Nsa
    and Wsa are all-reduced (SUM) before the policy is formed.  The ranks' trees are made
    different by giving rank r (40 + 5 r) simulations; the merged counts must sum to the total."""
    from visapi_b200 import runtime
    from visapi_b200.game import BinPackGame
    from visapi_b200.puct import BatchedMCTS
    runtime.set_default_device(local_rank)
    h, w, n, B = 20, 20, 15, 64
    np.random.seed(7)
    games = [BinPackGame(h, w, n, device=local_rank) for _ in range(B)]
    tiles = np.array([[[int(t[0]), int(t[1])] for t in g._base_board.tiles] for g in games])
    sims = 40 + 5 * rank

    def predict_dev(gr, tl):
        f = gr.sum(dim=(1, 2)).to(torch.int64)
        a = torch.arange(2 * n, dtype=torch.int64, device=gr.device)
        return ((a[None, :] * 7 + f[:, None] * 3) % 11 + 1).to(torch.float32), (f % 8).to(torch.float32) / 8.0
    t = BatchedMCTS(h, w, 2 * n, None, sims, 1.0, B, device=local_rank)
    t.set_roots(np.zeros((B, h, w)), tiles)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    t.simulate_on_device(predict_dev)
    own = t._tree.root_counts()[0].sum(axis=1)
    probs, q = t.action_probs_root_parallel(temp=1)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    own_all = torch.tensor([float(own.sum())], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(own_all)
    return {"trees": B, "sims_this_rank": sims, "root_visits_this_rank": int(own.sum()),
            "root_visits_all_ranks": int(own_all.item()),
            "policy_rows_sum_to_one": bool(np.allclose(probs.sum(axis=1), 1.0)), "seconds": dt,
            "collective": "all_reduce SUM of root Nsa (int32) and Wsa (float64), NCCL" if world > 1 else "none (1 rank)"}


def selfplay_record(local_rank, torch, dist, world, rank, games=1024, sims=50):
    """Config 5: BinPackGame(20,20,15), numMCTSSims = 50, cpuct = 1, fixed-weights PyTorch net; the
    ranks are replicas (each plays its own games)."""
    from visapi_b200 import runtime
    from visapi_b200.coach import BatchedSelfPlay
    from visapi_b200.game import BinPackGame
    from visapi_b200.nnet import TorchBinPackNet
    runtime.set_default_device(local_rank)
    np.random.seed(rank)
    game = BinPackGame(20, 20, 15, device=local_rank)
    net = TorchBinPackNet(game, device="cuda:%d" % local_rank)
    args = {"numMCTSSims": sims, "cpuct": 1.0, "tempThreshold": 15}
    sp = BatchedSelfPlay(20, 20, 15, net, args, n_games=games, device=local_rank, seed=rank)
    sp.play()                                            # warm-up (cudnn autotune, allocations)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    examples, reward, searches = sp.play()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    agg = torch.tensor([float(searches), float(len(examples))], dtype=torch.float64, device="cuda")
    tt = torch.tensor([dt], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(agg)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    return {"config": "BinPackGame(20,20,15), numMCTSSims=%d, cpuct=1, %d games per GPU in lock step" % (sims, games),
            "searches_per_s": float(agg[0].item()) / float(tt.item()), "seconds": float(tt.item()),
            "examples": int(agg[1].item()), "mean_leaf_batch": sp.search.leaf_evals / max(sp.search.leaf_batches, 1),
            "leaf_path": "device-resident leaves, torch net on the engine's stream",
            "scaling": "replicas (no collective on the search path)"}


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

    def reduce_max(x):
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def reduce_sum(values):
        t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return [float(x) for x in t.tolist()]

    problems = datasets.load_problem_set(a.set)[:a.n_problems]
    P = len(problems)
    strat = 0 if a.strategy == "max_depth" else 1
    mine = list(range(rank, P, world))                       # this rank's shard of THE set
    shard = datasets.batch_arrays([problems[i] for i in mine])
    shard_streams = np.asarray(mine, dtype=np.uint32)        # Philox stream = index in the set
    full = datasets.batch_arrays(problems)

    stream = torch.cuda.current_stream()
    eng = Engine(local_rank, stream=stream.cuda_stream)
    search = eng.search(shard["rows"], shard["cols"], shard["tiles"], shard["n_entries"], a.n_sim,
                        strat, a.seed, shard_streams)
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
    ms_rank = ev0.elapsed_time(ev1)
    clocks = sampler.stop(t_wall0, t_wall0 + wall)
    res = search.results()
    search_form = search.last_run_form()
    resident = search_form == "resident"
    sky_kernels = search.last_run_skyline()

    ms_max = reduce_max(ms_rank)
    ms_min = -reduce_max(-ms_rank)
    rollouts, placements, rollouts_exec, placements_exec, solved, launches_all = reduce_sum(
        [res["rollouts"].sum(), res["rollout_placements"].sum(), res["rollouts_executed"].sum(),
         res["placements_executed"].sum(), res["solution_found"].sum(), launches])
    sec_per_step = ms_max / 1e3 / a.steps
    value = rollouts / sec_per_step

    # ---- correctness gates -------------------------------------------------------------
    # (1) every reported solution is replayed on the host: first free cell, fits, fills the board
    #     exactly (visapi_b200/verify.py restates engine/solution_checker.py's placement rules; the
    #     reference file itself is not on the GPU box);
    # (2) N > 1: this rank's rows equal the same rows of the whole set searched on ONE GPU.
    verified = None
    if not a.no_verify:
        bad = []
        for k, i in enumerate(mine):
            if res["solution_found"][k]:
                p = problems[i]
                ok, why = replay_solution(p.rows, p.cols, p.tiles, res["order"][k, :res["n_order"][k]].tolist())
                if not ok:
                    bad.append((i, why))
        if bad:
            raise SystemExit("INVALID SOLUTIONS reported by the device: %r" % bad[:5])
        verified = int(reduce_sum([res["solution_found"].sum()])[0])
    weak = None
    shard_equals_full = None
    full_streams = np.arange(P, dtype=np.uint32)
    if world > 1 and not (a.no_weak and a.no_verify):
        fs = eng.search(full["rows"], full["cols"], full["tiles"], full["n_entries"], a.n_sim, strat, a.seed,
                        full_streams)
        fs.set_mode(a.mode)
        for _ in range(2):
            fs.reset()
            fs.run()
        barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_weak = max(2, min(a.steps, 5))
        e0.record(stream)
        for _ in range(n_weak):
            flush_buf.zero_()
            fs.reset()
            fs.run()
        e1.record(stream)
        torch.cuda.synchronize()
        barrier()
        rf = fs.results()
        same = all(np.array_equal(res[k], rf[k][mine]) for k in
                   ("depth", "solution_found", "n_tiles_placed", "rollouts", "rollout_placements", "n_order", "order"))
        shard_equals_full = bool(-reduce_max(-(1.0 if same else 0.0)) == 1.0)
        if not shard_equals_full:
            raise SystemExit("a rank's shard differs from the single-GPU search of the whole set")
        t_weak = reduce_max(e0.elapsed_time(e1)) / 1e3 / n_weak
        weak = {"rollouts_per_s": reduce_sum([rf["rollouts"].sum()])[0] / t_weak, "ms_per_step": 1e3 * t_weak,
                "problems_per_step": P * world,
                "note": "every rank searches the whole set (same Philox streams): replica throughput, "
                        "not the stated config"}
        fs.close()

    # ---- per-kernel time (CUDA events on the launching stream, classes serialised) -----
    search.set_profiling(True)
    step()
    rollout_ms, advance_ms, n_depths = search.profile()
    res_p = search.results()
    search.set_profiling(False)
    lane_kernel = search.info()["lane_kernel"]
    my_problems = [problems[i] for i in mine]
    if lane_kernel:
        keys = [lane_class_key(p.rows, p.cols, 2 * len(p.tiles)) for p in my_problems]
        per_placement = lambda k: model_lane_inst_per_placement(k) / 32.0   # noqa: E731
        min_ops = lambda k: min_ops_lane_per_placement(k) / 32.0            # noqa: E731
    else:
        keys = [shape_class_key(p.rows, p.cols, 2 * len(p.tiles)) for p in my_problems]
        per_placement = model_inst_per_placement
        min_ops = None
    by_class = {}
    for k, pe in zip(keys, res_p["placements_executed"].tolist()):
        by_class[k] = by_class.get(k, 0) + pe
    algo_inst = sum(pe * per_placement(k) for k, pe in by_class.items())
    min_inst = sum(pe * min_ops(k) for k, pe in by_class.items()) if min_ops else None
    info = eng.device_info()
    peaks = {}
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            peaks = json.load(f)
    except (OSError, ValueError):
        pass
    sm_max_mhz = float(peaks.get("sm_max_mhz") or clocks.get("sm_max_mhz") or 1965.0)
    issue_peak = info["sm_count"] * 4 * sm_max_mhz * 1e6
    # This rank's algorithmic work / this rank's step time (the class launches overlap on their
    # own streams in the timed region, so a rollout kernel's duration there is not separable;
    # every rollout kernel of a step runs inside the step, hence a lower bound of their rate).
    sec_rank = ms_rank / 1e3 / a.steps
    achieved = algo_inst / sec_rank
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    n_launches = 1 if resident else n_depths * search.info()["n_classes"]
    alu_only, alu_fma = eng.issue_microbench()
    n_tasks = float(res_p["rollouts_executed"].sum()) / 32.0
    algo_bytes = n_tasks * (128 + 8) + float(res_p["depth"].sum() + len(mine)) * 2400.0
    roofline = {
        "bound": "issue",
        "kernel": ("fs_search_resident_kernel" if resident else
                   (("fs_rollout_sky_kernel" if sky_kernels else "fs_rollout_lanes_kernel")
                    if lane_kernel else "fs_rollout_kernel")),
        "achieved": achieved / 1e9, "peak": issue_peak / 1e9, "unit": "Gwarp-inst/s",
        "frac": achieved / issue_peak,
        "frac_min_ops": (min_inst / sec_rank / issue_peak) if min_inst else None,
        "frac_is": "rank 0's share of the step: frozen work model A(class) (DESIGN.md 5) x placements "
                   "executed / step time / issue peak; frac_min_ops uses the implementation-independent "
                   "minimum operation count instead of A",
        "traffic": MEASURED_LANE_DRAM_BYTES_PER_LAUNCH if lane_kernel else 98560.0,
        "traffic_unit": "DRAM bytes per launch (ncu --set full, profiles/)",
        "algorithmic_bytes_per_launch": algo_bytes / max(n_launches, 1),
        "peak_source": "SMs x 4 schedulers x sm_max_mhz (%s)" %
                       ("MEASURED_PEAKS.json" if "sm_max_mhz" in peaks else "nvidia-smi"),
        "measured_issue_gwarp_inst_s": {"lop3_iadd3_alu_pipe_only": alu_only,
                                        "lop3_imad_alu_plus_fma": alu_fma},
        "ncu_pipe_pct_of_peak": MEASURED_LANE_PIPE_PCT if lane_kernel else None,
        "model": ("thread instructions per placement / 32 (lane-per-rollout)" if lane_kernel
                  else "warp instructions per placement (warp-per-board)"),
        "serialised_pass": {
            "note": "profiling pass: shape classes serialised, CUDA events around every launch; its "
                    "launches each pay their own ramp and tail, so their sum exceeds the step time of "
                    "the timed region, where the classes' launches overlap -- not evidence about the "
                    "timed region, only the rollout / advance split",
            "serialised_rollout_ms": rollout_ms, "serialised_advance_ms": advance_ms,
            "rollout_share": rollout_ms / max(rollout_ms + advance_ms, 1e-9),
            "frac": algo_inst / (rollout_ms / 1e3) / issue_peak},
        "ncu_in_flight_share": {"rollout_kernels": 0.919, "advance_kernels": 0.075,
                                "source": "profiles/r2/launches_r2_final.csv (launch list of this command)"},
        "launches_per_step": ({"resident": 1, "seed": 1} if resident
                              else {"rollout": n_launches, "advance": n_launches}),
        "placements_executed_per_step_this_rank": float(res_p["placements_executed"].sum()),
        "algorithmic_inst_per_placement": {k: per_placement(k) for k in by_class},
        "min_ops_inst_per_placement": ({k: min_ops(k) for k in by_class} if min_ops else None),
        "ncu_inst_per_placement": ({k: MEASURED_LANE_WARP_INST.get(k) for k in by_class}
                                   if lane_kernel else
                                   {k: MEASURED_INST_PER_PLACEMENT_V1.get(k) for k in by_class}),
        "hbm": {"algorithmic_gbs": algo_bytes / sec_rank / 1e9, "peak_gbs": hbm_peak,
                "frac": algo_bytes / sec_rank / 1e9 / hbm_peak},
    }

    # ---- end to end through the one-shot C-ABI call with host buffers --------------
    e2e = None
    if not a.no_e2e:
        h2d = (shard["tiles"].nbytes + shard["rows"].nbytes + shard["cols"].nbytes +
               shard["n_entries"].nbytes + shard_streams.nbytes)
        d2h = len(mine) * (56 + (shard["tiles"].shape[1] // 2 + 2) * 2 + (shard["tiles"].shape[1] // 2) * 4)
        for _ in range(2):
            eng.flat_search(shard["rows"], shard["cols"], shard["tiles"], shard["n_entries"], a.n_sim,
                            strat, a.seed, shard_streams)
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        e2e_rollouts = 0.0
        for _ in range(a.steps):
            flush_buf.zero_()
            r = eng.flat_search(shard["rows"], shard["cols"], shard["tiles"], shard["n_entries"],
                                a.n_sim, strat, a.seed, shard_streams)
            e2e_rollouts += float(r["rollouts"].sum())
        torch.cuda.synchronize()
        dt = reduce_max(time.perf_counter() - t0)
        total = reduce_sum([e2e_rollouts, h2d, d2h])
        e2e = {"value": total[0] / dt, "unit": "rollouts/s",
               "h2d_bytes_per_step": int(total[1]), "d2h_bytes_per_step": int(total[2]),
               "ms_per_step": 1e3 * dt / a.steps,
               "api": "bp_flat_search on every rank's shard (create + upload + all depths + download + destroy)"}

    # ---- sub-records ---------------------------------------------------------------
    extras = {}
    if not a.no_extras:
        for name, fn in (("root_parallel", lambda: root_parallel_record(eng, torch, dist, world, rank, dev)),
                         ("puct_root_parallel", lambda: puct_root_parallel_record(local_rank, torch, dist, world, rank, dev)),
                         ("selfplay", lambda: selfplay_record(local_rank, torch, dist, world, rank))):
            try:
                extras[name] = fn()
            except Exception as exc:            # reported, never allowed to lose the headline
                if world > 1:
                    raise
                extras[name] = {"error": "%s: %s" % (type(exc).__name__, str(exc)[:300])}

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
            "ms_per_step": 1e3 * sec_per_step, "higher_is_better": True,
            "scaling": "strong" if world > 1 else "weak",
            "vs_baseline": None, "dtype": "u32", "data": "reference problem set (bundled npz)",
            "config": workload_config(a),
            "rollouts_per_step": rollouts, "placements_per_step": placements,
            "rollouts_executed_per_step": rollouts_exec,
            "placements_executed_per_step": placements_exec,
            "placements_per_s": placements / sec_per_step,
            "rollouts_executed_per_s": rollouts_exec / sec_per_step,
            "solved": int(solved), "solutions_verified": verified,
            "verification": "every reported solution replayed by visapi_b200/verify.py (the repo's "
                            "restatement of engine/solution_checker.py's placement rules; the reference "
                            "file is not on the GPU box)" + ("; every rank's shard equals the same rows of "
                            "the whole set searched on one GPU" if shard_equals_full else ""),
            "shard_equals_single_gpu": shard_equals_full,
            "problems_per_step": P, "wall_s_per_step": wall / a.steps,
            "ms_per_step_fastest_rank": ms_min / a.steps, "search_mode": search_form,
            "clocks": clocks, "gpu_launches": int(launches_all),
            "roofline": roofline, "e2e": e2e, "weak": weak, "cpu_baseline": cpu,
        }
        line.update(extras)
        print(json.dumps(line))
    search.close()
    eng.close()
    if world > 1:
        barrier()
        dist.destroy_process_group()


def main():
    a = parse_args()
    if a.impl == "reference":
        reference_arm(a)
    else:
        our_arm(a)


if __name__ == "__main__":
    main()
