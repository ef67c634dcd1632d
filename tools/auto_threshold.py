"""Stepped versus resident search for small batches of problems_g (N = 1000): where AUTO switches."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visapi_b200 import Engine, datasets
probs = datasets.load_problem_set("g")
eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
for n_sim in (1000, 100):
    for P in (1, 2, 4, 8, 16, 32, 64, 125, 250):
        sub = datasets.batch_arrays(probs[:P])
        row = {"P": P, "n_sim": n_sim}
        for mode in ("stepped", "resident"):
            s = eng.search(sub["rows"], sub["cols"], sub["tiles"], sub["n_entries"], n_sim, 0, 123)
            s.set_mode(mode)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            ts = []
            for i in range(6):
                s.reset(); e0.record(); s.run(); e1.record(); torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            row[mode] = round(min(ts[2:]), 4)
            s.close()
        print(json.dumps(row), flush=True)
