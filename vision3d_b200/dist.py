"""Frame sharding across GPUs. Frames are independent (no cross-frame op anywhere in the
reference: BatchNorm in eval, NMS groups never span frames -- detector/proposal.py:32-39), so rank r owns a
contiguous block of frames and runs the whole path locally; the ONE exchange step is an all-gather of
the final, statically padded detections. The reference has no multi-GPU code at all (training.md:6).

One process per GPU, torch.distributed for the plumbing (NCCL over NVLink on the GPU box; gloo in the
CPU tests). The collective moves <= 11 floats x (B_local * n_cls * 100 + 1) rows per rank (106 KB at
config 5): latency bound, so it is a single ncclAllGather on the compute stream, no custom kernel.
"""
import os

import numpy as np
import torch
import torch.distributed as dist


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(
        os.environ.get("LOCAL_RANK", "0"))


def init(backend=None):
    """Initialise the default process group from the torchrun environment (MASTER_ADDR/PORT, RANK, ...)."""
    rank, world, local = env_rank_world()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def frame_range(n_frames, rank, world):
    """Contiguous block of frames owned by `rank` (config 5: rank r takes [8r, 8r+8) of 64)."""
    per = -(-n_frames // world)
    lo = min(rank * per, n_frames)
    return lo, min(lo + per, n_frames)


def gather_results(result, out=None):
    """All-gather the packed, statically padded per-rank result (rows, 11) -> (world, rows, 11).
    Asynchronous on the current stream with NCCL; no host synchronisation."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return result.unsqueeze(0)
    if out is None:
        out = result.new_empty((world,) + tuple(result.shape))
    if dist.get_backend() == "nccl":
        dist.all_gather_into_tensor(out.view(-1), result.contiguous().view(-1))
    else:  # gloo (CPU tests of the sharding logic)
        dist.all_gather(list(out.unbind(0)), result.contiguous())
    return out


def unpack_global(gathered, frames_per_rank):
    """Host side: (world, N+1, 11) numpy -> (boxes, batch_idx, class_idx, scores) with GLOBAL frame ids
    (local batch index + rank * frames_per_rank), ranks concatenated in order."""
    boxes, bidx, cidx, scores = [], [], [], []
    for r in range(gathered.shape[0]):
        rows = gathered[r, :-1]
        m = rows[:, 10] > 0
        boxes.append(rows[m, :7])
        bidx.append(rows[m, 8].astype(np.int64) + r * frames_per_rank)
        cidx.append(rows[m, 9].astype(np.int64))
        scores.append(rows[m, 7])
    return np.concatenate(boxes), np.concatenate(bidx), np.concatenate(cidx), np.concatenate(scores)
