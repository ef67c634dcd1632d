"""Device-side timeline of the launch-per-depth search of one GPU's share of problems_g.

Needs the debug library (every warp stamps %globaltimer at entry and exit; not shipped):

    nvcc ... -DBP_TIMELINE -c visapi_b200/csrc/flat_search.cu -o /tmp/flat_search_tl.o
    nvcc -shared -o visapi_b200/librectpack_b200_timeline.so /tmp/flat_search_tl.o \
         visapi_b200/csrc/build/{api_basic,puct,resident_k22,resident_k24,resident_k42,resident_k44}.o
    python tools/timeline.py [--world 8] [--n-sim 1000] > profiles/r2/timeline_shard.txt

Per launch: first warp start, last warp end (µs from the first start of the run), mean warp
lifetime and warps; then per chain the time between launches (nothing of this chain on the GPU),
and the share of the run in which fewer than `--thin` warps of any rollout kernel were alive.
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--world", type=int, default=8)
    ap.add_argument("--rank", type=int, default=0)
    ap.add_argument("--n-sim", type=int, default=1000)
    ap.add_argument("--set", default="g")
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "timeline.jsonl"))
    ap.add_argument("--mode", default="stepped", help="stepped | hybrid")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    os.environ.setdefault("BP_LIB", os.path.join(ROOT, "visapi_b200", "librectpack_b200_timeline.so"))
    os.environ["BP_TIMELINE_OUT"] = a.out
    os.makedirs(os.path.dirname(a.out), exist_ok=True)
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
    s.set_mode(a.mode)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(5):
        s.reset()
        e0.record(stream)
        s.run()
        e1.record(stream)
        torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    s.results()                                  # writes the timeline of the last run
    if s.last_run_form() != "stepped":
        st = s.resident_stats()
        print("resident part: " + json.dumps({k: (round(v, 4) if isinstance(v, float) else v) for k, v in st.items()
                                              if not isinstance(v, dict)}))
        print("  advance cycles: " + json.dumps({k: int(v) for k, v in st["advance_cycles"].items()}))
    s.close()
    eng.close()

    chains, launches = [], []
    for line in open(a.out):
        rec = json.loads(line)
        (launches if "kind" in rec else chains).append(rec)
    end = max(r["end_us"] for r in launches)
    print("search %.3f ms by CUDA events; first warp start to last warp end %.1f us; %d launches with work"
          % (ms, end, len(launches)))
    for c in chains:
        mine_l = sorted((r for r in launches if r["chain"] == c["chain"]), key=lambda r: (r["depth"], r["kind"] != "rollout"))
        roll = [r for r in mine_l if r["kind"] == "rollout"]
        adv = [r for r in mine_l if r["kind"] == "advance"]
        gaps = [b["start_us"] - a_["end_us"] for a_, b in zip(mine_l, mine_l[1:])]
        print("chain %d: %d problems, %d depths | rollout launches %d, span sum %.0f us, mean warp life %.1f us | "
              "advance launches %d, span sum %.0f us | between launches %.0f us (mean %.1f) | ends at %.0f us"
              % (c["chain"], c["problems"], c["depths"], len(roll), sum(r["end_us"] - r["start_us"] for r in roll),
                 sum(r["warp_us"] for r in roll) / max(1, sum(r["warps"] for r in roll)),
                 len(adv), sum(r["end_us"] - r["start_us"] for r in adv), sum(gaps), sum(gaps) / max(1, len(gaps)),
                 mine_l[-1]["end_us"] if mine_l else 0))
    # how full is the GPU: integral of live rollout warps, approximated per launch as warp_us spread
    # evenly over the launch's span
    step = 1.0
    n = int(end / step) + 1
    live = np.zeros(n)
    for r in launches:
        if r["kind"] != "rollout":
            continue
        i0, i1 = int(r["start_us"] / step), max(int(r["end_us"] / step), int(r["start_us"] / step) + 1)
        live[i0:i1] += r["warp_us"] / max(r["end_us"] - r["start_us"], step)
    print("live rollout warps (launch average spread over its span): mean %.0f, share of the run below 1000: %.1f %%, "
          "below 2000: %.1f %%, below 3000: %.1f %%" % (live.mean(), 100 * (live < 1000).mean(),
                                                      100 * (live < 2000).mean(), 100 * (live < 3000).mean()))
    total_warp_us = sum(r["warp_us"] for r in launches if r["kind"] == "rollout")
    print("sum of rollout warp lifetimes %.0f warp-us = %.0f warps busy for the whole run" % (total_warp_us, total_warp_us / end))
    if a.verbose:
        for r in sorted(launches, key=lambda r: r["start_us"]):
            print("%8.1f %8.1f  d=%2d chain=%d %-7s warps=%5d mean_life=%6.1f" % (
                r["start_us"], r["end_us"], r["depth"], r["chain"], r["kind"], r["warps"], r["warp_us"] / r["warps"]))


if __name__ == "__main__":
    main()
