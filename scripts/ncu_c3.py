"""Run a few steps of the PV-RCNN keypoint stage (config C3) so that ncu can see every launch.
    ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file c3_launches.csv python scripts/ncu_c3.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision3d_b200 import pvrcnn, synth  # noqa: E402

B, n = 8, 100
dev = torch.device("cuda:0")
cfg = pvrcnn.PVRCNNConfig()
model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0)
stage = pvrcnn.KeypointStage(model, B, 16384, n, dev)
clouds = synth.make_batch(0, B, 16384)
props = pvrcnn.make_proposals(clouds, n, 0)
grid = pvrcnn.sample_gridpoints(torch.from_numpy(props), pvrcnn.make_grid_noise(B, n, 16, 0)).reshape(B, -1, 3)
stage.load(clouds, grid)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    torch.cuda.nvtx.range_push("c3_step")
    stage.step()
    torch.cuda.synchronize()
    torch.cuda.nvtx.range_pop()
print("pooled", tuple(stage.pooled.shape))
