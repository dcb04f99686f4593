"""FPS configuration sweep (experiment): times v3d_fps_keypoints at the C3 shape for every V3D_FPS_CFG and checks that
all configurations return the same indices.
Historical record of the sweep behind profiles/r02_fps_sweep.json: the V3D_FPS_CFG switch existed only in the experiment
build; the product dispatches to the winner (4 warps x 8-CTA cluster) and this script now times that one configuration."""
import os, sys, json, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision3d_b200 import ops

dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(0)
B, N, m = 8, 16384, 2048
pts = (torch.rand((B, N, 4), generator=g) * torch.tensor([70.4, 80.0, 4.0, 1.0]) + torch.tensor([0.0, -40.0, -3.0, 0.0])).to(dev)
names = {0: "16w x cl8", 1: "8w x cl8", 2: "4w x cl8", 3: "4w x cl16", 4: "8w x cl16", 5: "2w x cl16", 6: "barrier kernel"}
ref, res = None, {}
for c in [6, 0, 1, 2, 3, 4, 5]:
    os.environ["V3D_FPS_CFG"] = str(c)
    try:
        idx, kp = ops.fps_keypoints(pts, m)
        torch.cuda.synchronize()
        if ref is None:
            ref = idx.clone()
        same = bool(torch.equal(idx, ref))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.fps_keypoints(pts, m)
        e1.record(); torch.cuda.synchronize()
        res[names[c]] = {"ms": round(e0.elapsed_time(e1) / 5, 4), "same": same}
    except Exception as ex:  # a configuration the device refuses (cluster 16 placement) is reported, not fatal
        res[names[c]] = {"error": str(ex)[:200]}
    print(names[c], res[names[c]], flush=True)
# small / ragged shapes through the default configuration
os.environ["V3D_FPS_CFG"] = "6"
for (b, n, mm) in [(3, 1000, 64), (2, 4097, 300), (1, 33, 33), (2, 16384, 16)]:
    p = torch.rand((b, n, 3), generator=g).to(dev)
    os.environ["V3D_FPS_CFG"] = "6"; want = ops.fps_keypoints(p, mm)[0]
    for c in [0, 1, 2, 3]:
        os.environ["V3D_FPS_CFG"] = str(c)
        try:
            got = ops.fps_keypoints(p, mm)[0]
            print((b, n, mm), names[c], bool(torch.equal(got, want)), flush=True)
        except Exception as ex:
            print((b, n, mm), names[c], "error", str(ex)[:100], flush=True)
json.dump(res, open("gpurun_out/r02_fps_sweep.json", "w"), indent=1)
