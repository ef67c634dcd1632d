"""One GPU's share of an 8-way split of problems_g (and the full set): ms per search for
several work-unit sizes / launch structures, CUDA events on the launching stream.

    python tools/shard_sweep.py [--n-sim 1000] [--world 8] [--reps 10]
"""
import argparse
import json
import os
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def child(a):
    import numpy as np
    import torch

    from visapi_b200 import Engine, datasets
    problems = datasets.load_problem_set(a.set)[:1000]
    mine = list(range(a.rank, len(problems), a.world))
    sub = datasets.batch_arrays([problems[i] for i in mine])
    stream = torch.cuda.current_stream()
    eng = Engine(0, stream=stream.cuda_stream)
    s = eng.search(sub["rows"], sub["cols"], sub["tiles"], sub["n_entries"], a.n_sim, 0, 123,
                   np.asarray(mine, dtype=np.uint32))
    if a.mode:
        s.set_mode(a.mode)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for _ in range(3):
        s.reset()
        s.run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    times = []
    for _ in range(a.reps):
        flush.zero_()
        s.reset()
        e0.record(stream)
        s.run()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    res = s.results()
    out = {"ms_min": min(times), "ms_med": float(np.median(times)), "problems": len(mine),
           "rollouts": int(res["rollouts"].sum()), "solved": int(res["solution_found"].sum()),
           "placements_executed": int(res["placements_executed"].sum()),
           "digest": int(np.bitwise_xor.reduce(res["rollouts"].astype(np.uint64) * np.uint64(2654435761)
                                               + res["depth"].astype(np.uint64)))}
    if s.last_run_form() != "stepped":
        out["resident"] = {k: (round(v, 4) if isinstance(v, float) else v) for k, v in s.resident_stats().items()}
        out["resident"]["advance_cycles"] = {k: int(v) for k, v in out["resident"]["advance_cycles"].items()}
    if a.dump:
        np.save(a.dump, np.stack([res["placements_executed"], res["rollouts_executed"],
                                  res["depth"], res["solution_found"]]))
    print(json.dumps(out))
    s.close()
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--child", action="store_true")
    ap.add_argument("--set", default="g")
    ap.add_argument("--n-sim", type=int, default=1000)
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--mode", default="")
    ap.add_argument("--dump", default="")
    ap.add_argument("--envs", default="", help="';'-separated list of 'K=V,K=V' settings to compare")
    a = ap.parse_args()
    if a.child:
        return child(a)
    envs = [e for e in a.envs.split(";")] if a.envs else [""]
    for world in sorted(set([a.world, 1])):
        for setting in envs:
            env = dict(os.environ)
            for kv in filter(None, setting.split(",")):
                k, v = kv.split("=")
                env[k] = v
            cmd = [sys.executable, __file__, "--child", "--set", a.set, "--n-sim", str(a.n_sim), "--world",
                   str(world), "--rank", str(a.rank), "--reps", str(a.reps)]
            if a.mode:
                cmd += ["--mode", a.mode]
            if a.dump and world == 1:
                cmd += ["--dump", a.dump]
            out = subprocess.run(cmd, env=env, capture_output=True, text=True)
            line = out.stdout.strip().splitlines()[-1] if out.stdout.strip() else out.stderr[-400:]
            print("world=%d %-40s %s" % (world, setting or "(default)", line), flush=True)


if __name__ == "__main__":
    main()
