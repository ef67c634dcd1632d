"""ms per bp_search_reset (memsets, copies of the initial states, the first advance of every class)."""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from visapi_b200 import Engine, datasets
probs = datasets.load_problem_set("g")
eng = Engine(0, stream=torch.cuda.current_stream().cuda_stream)
for P in (1, 125, 1000):
    sub = datasets.batch_arrays(probs[:P] if P < 1000 else probs)
    s = eng.search(sub["rows"], sub["cols"], sub["tiles"], sub["n_entries"], 1000, 0, 123)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        s.reset()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(20):
        s.reset()
    e1.record()
    torch.cuda.synchronize()
    print(json.dumps({"P": P, "reset_ms": e0.elapsed_time(e1) / 20}))
    s.close()
