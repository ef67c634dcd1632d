"""One 50-tile generated instance (config 4) on one GPU: ms per reset + search for each search form.

    python tools/c4_modes.py [--n-sim 5000 1000 100]
"""
import argparse
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tools.bench_root_parallel import instance  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n-sim", type=int, nargs="+", default=[5000, 1000, 100])
    ap.add_argument("--seeds", type=int, default=4)
    a = ap.parse_args()
    from visapi_b200 import Engine, datasets
    eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
    for n_sim in a.n_sim:
        for (rows, cols) in ((25, 25), (40, 40)):
            row = {"n_sim": n_sim, "board": "%dx%d" % (rows, cols)}
            for mode in ("stepped", "resident", "auto"):
                tot = 0.0
                for seed in range(a.seeds):
                    batch = datasets.batch_arrays([instance(50, rows, cols, seed)])
                    s = eng.search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], n_sim, 0, 123,
                                   np.array([seed], np.uint32))
                    s.set_mode(mode)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    ts = []
                    for _ in range(6):
                        e0.record(); s.reset(); s.run(); e1.record(); torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                    tot += min(ts[2:])
                    form = s.last_run_form()
                    s.close()
                row[mode] = round(tot / a.seeds, 4)
                if mode == "auto":
                    row["auto_form"] = form
            print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
