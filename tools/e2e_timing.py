import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["BP_TIMING"] = "1"
import numpy as np
from visapi_b200 import Engine, datasets
probs = datasets.load_problem_set("g")
batch = datasets.batch_arrays(probs)
eng = Engine(0)
for i in range(4):
    t0 = time.perf_counter()
    r = eng.flat_search(batch["rows"], batch["cols"], batch["tiles"], batch["n_entries"], 1000, 0, 123)
    print("python call %.0f us" % ((time.perf_counter() - t0) * 1e6), flush=True)
