"""Host-side cost of one TrainEngine step (cProfile on the box): where the Python time of ~3 400 launches goes."""
import cProfile
import pstats
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "multimodal-auv_b200"))
sys.path.insert(0, str(ROOT))
import torch

import bench
from mauv.train_engine import TrainEngine

B, S = 8, 30
model = bench.build_model_cpu().cuda().train()
xs = [x.cuda() for x in bench.synthetic_inputs(B)]
labels = torch.randint(0, 7, (B,), device="cuda")
eng = TrainEngine(model)
eng.flatten_grads()
for _ in range(2):
    eng.zero_grad()
    eng.step(xs, labels, S, 1e-6)
torch.cuda.synchronize()
import time
t0 = time.perf_counter()
eng.zero_grad()
eng.step(xs, labels, S, 1e-6)
t1 = time.perf_counter()
torch.cuda.synchronize()
t2 = time.perf_counter()
print(f"host enqueue {1e3 * (t1 - t0):.1f} ms, until GPU done {1e3 * (t2 - t0):.1f} ms")
pr = cProfile.Profile()
pr.enable()
eng.zero_grad()
eng.step(xs, labels, S, 1e-6)
pr.disable()
torch.cuda.synchronize()
pstats.Stats(pr).sort_stats("tottime").print_stats(18)
