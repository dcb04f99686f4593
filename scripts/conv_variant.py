"""Per-op device time of the sparse-conv layers of the t16 workload for the current V3D_TC_* environment
(one process per setting: the knobs are read once). Usage: V3D_TC_FETCH=2 python scripts/conv_variant.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from vision3d_b200 import second, synth  # noqa: E402

dev = torch.device("cuda:0")
eng, model, cfg = second.make_bench_engine("t16", 16, dev, use_graph=False)
eng.load_host(synth.make_batch(0, 16))
eng.step_e2e()
torch.cuda.synchronize()
ops_t = eng.profile_ops(iters=5)
conv = [(n, round(t)) for n, t in ops_t if n.startswith(("subm_L", "sconv_L"))]
print("V3D_TC_FETCH=%s conv total %.0f us :" % (os.environ.get("V3D_TC_FETCH"), sum(t for _, t in conv)), conv,
      flush=True)
