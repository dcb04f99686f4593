"""world_size-2 gloo tests (CPU) of the frame-sharding / all-gather / unpack logic."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    from vision3d_b200 import dist as vdist
    r, w, _ = vdist.init(backend="gloo")
    assert (r, w) == (rank, world)
    frames = 8
    lo, hi = vdist.frame_range(frames, rank, world)
    n = 12
    # packed result of this rank: local batch idx, some valid rows, one trailing counter row
    res = torch.zeros((n + 1, 11))
    k = 3 + rank
    res[:k, 0] = torch.arange(k) + 100 * rank          # box x
    res[:k, 7] = 0.9 - 0.1 * torch.arange(k)            # score
    res[:k, 8] = torch.arange(k) % (hi - lo)            # LOCAL frame index
    res[:k, 9] = rank                                   # class
    res[:k, 10] = 1
    res[n, 0] = k
    g = vdist.gather_results(res)
    assert g.shape == (world, n + 1, 11)
    # in-place form bench.py uses: this rank's result lives in its slot of the gather buffer
    buf = torch.zeros((world, n + 1, 11))
    buf[rank].copy_(res)
    g2 = vdist.gather_results(buf[rank], buf)
    assert g2.data_ptr() == buf.data_ptr() and torch.equal(g2, g)
    boxes, bidx, cidx, scores = vdist.unpack_global(g.numpy(), hi - lo)
    ret[rank] = (lo, hi, boxes[:, 0].tolist(), bidx.tolist(), cidx.tolist())
    dist.barrier()
    dist.destroy_process_group()


def test_frame_sharding_and_gather_world2():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, 29591, ret), nprocs=world, join=True)
    assert ret[0][:2] == (0, 4) and ret[1][:2] == (4, 8)
    for r in range(world):  # every rank ends with the same global view
        lo, hi, xs, bidx, cidx = ret[r]
        assert xs == [0.0, 1.0, 2.0, 100.0, 101.0, 102.0, 103.0]
        assert bidx == [0, 1, 2, 4, 5, 6, 7]            # rank 1's local frames shifted by 4
        assert cidx == [0, 0, 0, 1, 1, 1, 1]


def test_frame_range_uneven():
    from vision3d_b200 import dist as vdist
    assert [vdist.frame_range(10, r, 4) for r in range(4)] == [(0, 3), (3, 6), (6, 9), (9, 10)]
    assert vdist.frame_range(2, 3, 4) == (2, 2)


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the GPU arm) needs no GPU: one JSON line
    with the contract keys, timed on the host cores."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "1"], capture_output=True, text=True, timeout=600, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["unit"] == "frames/s" and line["value"] > 0
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling", "dtype", "data",
              "config", "cpu_baseline", "e2e"):
        assert k in line, k
    assert line["e2e"]["h2d_bytes_per_step"] == 0 and line["cpu_baseline"]["cores"] >= 1
