"""Run a few eager (no CUDA graph) steps of the SECOND engine so that ncu can see every launch.
    ncu --metrics gpu__time_duration.sum --clock-control none -s 400 -c 200 --csv --log-file launches.csv \
        python scripts/ncu_step.py --steps 3
"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision3d_b200 import second, synth  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--steps", type=int, default=3)
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--simt", action="store_true")
ap.add_argument("--rpn", default="fused")
a = ap.parse_args()
dev = torch.device("cuda:0")
cfg = second.car_config()
model = second.init_for_benchmark(second.SecondB200(cfg), 0)
eng = second.SecondEngine(model, a.batch, a.batch * 16384, dev, use_graph=False, tensor_cores=not a.simt,
                          rpn_mode=a.rpn)
eng.load_host(synth.make_batch(0, a.batch))
with torch.no_grad():
    for _ in range(a.steps):
        torch.cuda.nvtx.range_push("step")
        eng.step_e2e()
        torch.cuda.synchronize()
        torch.cuda.nvtx.range_pop()
print("launches/step (v3d kernels):", eng.kernel_launches, eng.finalize()[0].shape)
