#!/usr/bin/env python
"""bench.py -- frames/s of the per-frame LiDAR hot path on synthetic KITTI-shape clouds.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

The JSON line's headline (`metric`/`value`/`e2e`) is workload **t16**: SECOND car-only, 16 frames PER GPU (the target
line "SECOND at batch 16"; weak scaling: frames are sharded, one all-gather of the final boxes per step when N > 1,
captured inside the CUDA graph). A "step" is one pass of the hot path over one batch: raw points -> voxelize+VFE ->
sparse 3-D backbone (rule books + 14 fused sparse convs) -> dense BEV -> RPN -> heads/top-k/decode -> rotated NMS.

  value : whole-job frames/s with the inputs already resident in HBM (one CUDA-graph replay per step); the K-step
          timed region (barrier + synchronize on both sides, CUDA events, max over ranks) is repeated `repeats` times
          and the MEDIAN repeat is reported (region >= 2 s in total, clocks sampled in-process through NVML)
  e2e   : same metric through the public call path with HOST buffers: pinned H2D of the raw points and D2H of the
          packed detections inside the timed region (packing the numpy clouds into the pinned buffer -- the
          `load_host` memcpy -- is outside it and said so in `config`)
  roofline / per_op : per-op device time measured live with CUDA events (eager pass, same stream), algorithmic
          bytes per SURVEY.md 8(d), peak from MEASURED_PEAKS.json
  cpu_baseline : the same path on the host CPU (oracle/second_cpu.py), bounded sample (N=1, rank 0)
  workloads.c5 : BASELINE config 5 -- SECOND 3-class, GLOBAL batch 64 sharded over the ranks (STRONG scaling: rank r
          takes frames [64 r / N, 64 (r+1) / N)), same timing rules, value = 64 K / t
  workloads.c3 : BASELINE config 3 -- PV-RCNN keypoint stage (FPS-2048 + 5-source ball query / grouping / MLP +
          RoI-grid pool), batch 8, per-op times with SURVEY 8(d)'s per-op figures, CPU baseline beside it (N=1 only)

--impl reference times the CPU path itself (the reference's own CPU NMS op from oracle/_ref when it was built; the
un-vendored spconv/pointnet2 parts are the oracle port) with all host threads.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "frames/sec on synthetic KITTI-shape clouds (SECOND car-only, per-GPU batch 16)"
PTS_PER_FRAME = 16384
C5_GLOBAL_BATCH = 64


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=16, help="frames per GPU per step (t16)")
    ap.add_argument("--repeats", type=int, default=0, help="timed regions per measurement (0 = enough for >= 2 s)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-c5", action="store_true")
    ap.add_argument("--no-c3", action="store_true")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--rpn", default="fused_nhwc", choices=["module", "fused", "fused_nhwc"])
    ap.add_argument("--simt", action="store_true", help="exact-fp32 SIMT sparse conv instead of tcgen05 bf16x3")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled every 10 ms DURING the timed regions, in-process through NVML
    (nvidia-ml-py); falls back to an `nvidia-smi -lms 20` child when NVML cannot be loaded."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index):
        self.sm, self.power, self.reasons, self.max_sm = [], [], set(), None
        self.stop_flag, self.thread, self.proc, self.source = False, None, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # torch's device index is relative to CUDA_VISIBLE_DEVICES
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = gpu_index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if gpu_index < len(ids) and ids[gpu_index].isdigit():
                    phys = int(ids[gpu_index])
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_sm = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.source = "nvml"
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
        except Exception:
            self._start_smi(gpu_index)

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in self.REASONS.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.01)

    def _start_smi(self, gpu_index):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(gpu_index), "--query-gpu=" + q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.source = "nvidia-smi"
            self.thread = threading.Thread(target=self._read_smi, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read_smi(self):
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.proc.stdout:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 7:
                continue
            try:
                self.sm.append(float(f[0]))
                self.max_sm = float(f[1])
                self.power.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[3:7]):
                if v == "Active":
                    self.reasons.add(n)

    def stop(self):
        self.stop_flag = True
        if self.proc is not None:
            self.proc.terminate()
        if self.thread is not None:
            self.thread.join(timeout=2)
        if not self.sm:
            return {"sm_mhz": None, "sm_max_mhz": self.max_sm, "samples": 0, "reasons": ["no clock samples"],
                    "source": self.source}
        sm = sorted(self.sm)
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": self.max_sm, "samples": len(sm),
                "power_w_max": round(max(self.power), 1) if self.power else None, "reasons": sorted(self.reasons),
                "source": self.source, "period_ms": 10 if self.source == "nvml" else 20}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), float(d.get("bf16_tflops_sustained", d["bf16_tflops"])), "measured"
    return 6650.0, 1590.0, "fallback"


# ------------------------------------------------------------------------------------------- CPU arm
def build_cpu_model(seed=0):
    import torch
    from vision3d_b200 import second
    cfg = second.car_config()
    torch.manual_seed(seed)
    model = second.init_for_benchmark(second.SecondB200(cfg), seed).eval()
    return cfg, model


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


def cpu_frames_per_s(frames, warm=1):
    """The hot path on host cores for `frames` single-frame steps. Returns (fps, info)."""
    import torch
    from oracle import second_cpu
    from vision3d_b200 import second, synth
    cfg, model = build_cpu_model()
    anchors = second.make_anchors(cfg)
    # give the CPU arm its best thread count (many-core boxes lose to oversubscription on these small GEMMs); thread
    # counts are tried in ASCENDING order and the search stops as soon as more threads stop helping
    cores = host_cores()
    best = (None, float("inf"))
    warm_cloud = [synth.make_cloud(900, PTS_PER_FRAME)]
    torch.set_num_threads(min(cores, 8))
    second_cpu.infer(model, warm_cloud, anchors)
    for nt in sorted({min(cores, 4), min(cores, 8), min(cores, 16), min(cores, 32), min(cores, 64), cores}):
        torch.set_num_threads(nt)
        t = time.perf_counter()
        second_cpu.infer(model, warm_cloud, anchors)
        dt = time.perf_counter() - t
        if dt < best[1]:
            best = (nt, dt)
        elif dt > 1.15 * best[1]:
            break
    torch.set_num_threads(best[0])
    for i in range(max(0, warm - 1)):
        second_cpu.infer(model, [synth.make_cloud(901 + i, PTS_PER_FRAME)], anchors)
    ts = []
    for i in range(frames):
        cloud = [synth.make_cloud(i, PTS_PER_FRAME)]
        t = time.perf_counter()
        second_cpu.infer(model, cloud, anchors)
        ts.append(time.perf_counter() - t)
    import oracle
    kind = "port+reference-nms" if oracle.ref_available("ref_C_cpu.so") else "port"
    return frames / sum(ts), dict(cores=torch.get_num_threads(), kind=kind, step_s=ts)


CPU_SAMPLE = ("%d single-frame steps of the same synthetic workload through oracle/second_cpu.py: rotated NMS = the "
              "reference's own compiled CPU op (oracle/_ref/ref_C_cpu.so) when present; voxelize / rule book = oracle "
              "C++ port (1 thread); sparse conv = per-offset gather / torch.mm / scatter port, RPN and head in torch "
              "on %d threads")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import oracle
    oracle.lib()
    t0 = time.perf_counter()
    fps, info = cpu_frames_per_s(args.steps, warm=max(1, min(args.warmup, 2)))
    ms = 1e3 / fps
    line = {
        "impl": "reference", "metric": METRIC, "value": round(fps, 4), "unit": "frames/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms, 2), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "t16: SECOND car-only (configs/second/car.yaml), KITTI-shape synthetic clouds, "
                               "%d pts/frame; CPU arm: each step = 1 frame (bounded sample of the batch-16 step)"
                               % PTS_PER_FRAME, "frames_per_step": 1},
        "cpu_baseline": {"value": round(fps, 4), "unit": "frames/s", "cores": info["cores"], "kind": info["kind"],
                         "sample": CPU_SAMPLE % (args.steps, info["cores"])},
        "e2e": {"value": round(fps, 4), "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "wall_s": round(time.perf_counter() - t0, 1),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------- GPU arm
def algorithmic(engine, name, counts, pairs):
    """Algorithmic HBM bytes / flops of one plan op."""
    B = engine.B
    if name.startswith("voxelize"):
        return 16 * counts["points"] + 100 * counts["rows"][0], 0
    if name.startswith("dense"):
        sh = engine.shapes[4]
        return counts["rows"][4] * (4 * 64 + 16) + 4 * B * 64 * sh[0] * sh[1] * sh[2], 0
    if name.startswith("subm_L") or name.startswith("sconv_L"):
        lv = int(name.split("_L")[1][0])
        cin, cout = [int(v) for v in name.rsplit("_", 1)[1].split("x")]
        if name.startswith("subm"):
            n_in = n_out = counts["rows"][lv]
            P, kv = pairs["subm"][lv], 27
        else:
            n_in, n_out = counts["rows"][lv], counts["rows"][lv + 1]
            P, kv = pairs["conv"][lv], engine.nbr_conv[lv].shape[0]
        return 4 * (n_in * cin + n_out * cout) + 8 * P + 4 * kv * cin * cout, 2 * P * cin * cout
    if name.startswith("rulebook_subm"):
        lv = int(name[-1])
        return 32 * counts["rows"][lv] + 8 * pairs["subm"][lv], 0
    if name.startswith("rulebook_conv"):
        lv = int(name[-1])
        return 16 * counts["rows"][lv] + 16 * counts["rows"][lv + 1] + 8 * pairs["conv"][lv], 0
    if name.startswith("site_table"):
        lv = int(name[-1])
        return 16 * counts["rows"][lv], 0
    if name.startswith("nms"):
        return 24 * engine.N + 8 * counts["kept"], 0
    return 0, 0


class Harness:
    """Timing rules shared by every workload: W >= 3 warm-up steps, K-step regions bracketed by barrier + synchronize,
    CUDA events on the launching stream, max over ranks, median over `repeats` regions."""

    def __init__(self, dev, world, steps, repeats):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.dev, self.world, self.steps, self.repeats = torch, dist, dev, world, steps, repeats

    def region(self, fn):
        torch, dist = self.torch, self.dist
        if self.world > 1:
            dist.barrier()
        torch.cuda.synchronize(self.dev)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for j in range(self.steps):
            fn(j)
        b.record()
        torch.cuda.synchronize(self.dev)
        if self.world > 1:
            dist.barrier()
        ms = torch.tensor([a.elapsed_time(b)], device=self.dev)
        if self.world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    def measure(self, fn, repeats=None):
        """-> (median ms per region, all region times). `repeats` None: enough regions for >= 2 s (5..50)."""
        first = self.region(fn)
        n = repeats if repeats else self.repeats
        if not n:
            n = max(5, min(50, int(2000.0 / max(first, 1e-3)) + 1))
        ts = sorted([first] + [self.region(fn) for _ in range(n - 1)])
        return ts[len(ts) // 2], ts


def second_workload(args, h, rank, world, dev, workload, frames_local, seed_base):
    """Build the benchmarked engine (second.make_bench_engine: the flags tests/test_gpu_bench_config.py parity-tests)
    and the rotating input sets. Returns (engine, step_device, step_e2e, sets)."""
    import torch
    from vision3d_b200 import dist as vdist
    from vision3d_b200 import second, synth
    overrides = {}
    if args.no_graph:
        overrides["use_graph"] = False
    if args.simt:
        overrides["tensor_cores"] = False
    if args.rpn != "fused_nhwc":
        overrides["rpn_mode"] = args.rpn
    gathered = None
    if world > 1:
        # ONE exchange step: all-gather of the padded detections, issued from inside the captured step. The engine
        # packs its result straight into this rank's slot of the gather buffer (in-place ncclAllGather).
        cfg_n = 1 if second.BENCH_WORKLOADS[workload]["cfg"] == "car" else 3
        rows = frames_local * cfg_n * 100 + 1
        gathered = torch.zeros((world, rows, 11), dtype=torch.float32, device=dev)
        overrides["result_out"] = gathered[rank]
        overrides["post_step"] = lambda: vdist.gather_results(gathered[rank], gathered)
        # communicator warm-up outside any capture
        vdist.gather_results(gathered[rank], gathered)
        torch.cuda.synchronize(dev)
    eng, model, cfg = second.make_bench_engine(workload, frames_local, dev, **overrides)
    n_sets = 4
    h_sets, d_sets = [], []
    for j in range(n_sets):
        clouds = synth.make_batch(seed_base + 1000 * j, frames_local, PTS_PER_FRAME)
        n = eng.load_host(clouds)
        h_sets.append((eng.h_points[:n].clone().pin_memory(), eng.h_off.clone().pin_memory(), n))
        d_sets.append((h_sets[-1][0].to(dev), h_sets[-1][1].to(dev), n))
    h_gathered = torch.empty(tuple(gathered.shape)).pin_memory() if world > 1 else None

    def step_device(j):
        p, o, n = d_sets[j % n_sets]
        eng.points[:n].copy_(p, non_blocking=True)
        eng.frame_off.copy_(o, non_blocking=True)
        eng.step_device()

    def step_e2e(j):
        p, o, n = h_sets[j % n_sets]
        eng.points[:n].copy_(p, non_blocking=True)
        eng.frame_off.copy_(o, non_blocking=True)
        eng.step_device()
        if world > 1:
            h_gathered.copy_(gathered, non_blocking=True)
        else:
            eng.h_result.copy_(eng.result, non_blocking=True)

    for j in range(max(args.warmup, 3)):
        step_e2e(j)
    torch.cuda.synchronize(dev)
    # every rank checks its own level counters (raises on a capacity overflow)
    eng.h_result.copy_(eng.result)
    eng._staged_points = h_sets[0][2]
    eng.finalize()
    return eng, step_device, step_e2e, h_sets, gathered


def per_op_table(eng, h_sets, hbm_peak, peak_kind):
    import torch
    per_op = eng.profile_ops(iters=5)
    torch.cuda.synchronize(eng.dev)
    rows = [int(eng.n_rows[lv].item()) for lv in range(5)]
    counts = dict(points=h_sets[0][2], rows=rows, kept=int(eng.count.item()))
    pairs = dict(subm=[int((eng.nbr_subm[lv][:, :rows[lv]] >= 0).sum().item()) for lv in range(4)],
                 conv=[int((eng.nbr_conv[lv][:, :rows[lv + 1]] >= 0).sum().item()) for lv in range(4)])
    step_us = sum(t for _, t in per_op)
    table = []
    for name, us in per_op:
        nbytes, flops = algorithmic(eng, name, counts, pairs)
        table.append({"op": name, "us": round(us, 2), "share": round(us / step_us, 4),
                      "alg_bytes": int(nbytes), "gbs": round(nbytes / us / 1e3, 1) if nbytes else None,
                      "tflops": round(flops / us / 1e6, 2) if flops else None})
    mine = [r for r in table if "(" not in r["op"] and r["alg_bytes"] and r["op"] != "pack_result"]
    groups = {}
    for r in mine:  # the "dominant kernel" = the v3d kernel family with the largest share of the step
        fam = "sparse_conv_fwd" if r["op"].startswith(("subm_L", "sconv_L")) else r["op"].split("_L")[0]
        g = groups.setdefault(fam, {"us": 0.0, "bytes": 0, "n": 0})
        g["us"] += r["us"]
        g["bytes"] += r["alg_bytes"]
        g["n"] += 1
    fam = max(groups, key=lambda k: groups[k]["us"])
    g = groups[fam]
    achieved = g["bytes"] / g["us"] / 1e3
    roofline = {"kernel": fam, "bound": "hbm", "achieved": round(achieved, 1), "peak": hbm_peak, "unit": "GB/s",
                "frac": round(achieved / hbm_peak, 4), "traffic": None, "peak_kind": peak_kind,
                "launches": g["n"], "avg_us_per_launch": round(g["us"] / g["n"], 2),
                "alg_bytes_per_launch": int(g["bytes"] / g["n"]), "share_of_step": round(g["us"] / step_us, 4)}
    # measured DRAM traffic per launch (dram__bytes_read.sum + dram__bytes_write.sum) comes from the committed
    # `ncu --set full` capture of the same workload (scripts/summarize_ncu.py -> profiles/traffic.json)
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        t = json.load(open(tpath)).get(fam)
        if t:
            roofline["traffic"] = int(t["dram_bytes_per_launch"])
            roofline["traffic_source"] = t.get("source")
    if fam == "sparse_conv_fwd":
        # what actually bounds this kernel (ncu counters): neither HBM nor the tensor pipe
        roofline["limiter"] = ("shared-memory crossbar: SS-operand MMAs move 104 KB per K=64 slot (A written + read, weight "
                               "image written + read) at 128 B/clk/SM = 832 clk per slot against 444 clk of MMAs; measured "
                               "0.92 wavefronts/clk before, 0.71 after absent neighbours stopped being copied "
                               "(profiles/r02y_conv_counters.md, profiles/r02x_conv_counters.md)")
    fams = {k: {"us": round(v["us"], 1), "gbs": round(v["bytes"] / v["us"] / 1e3, 1),
                "frac": round(v["bytes"] / v["us"] / 1e3 / hbm_peak, 4), "launches": v["n"]} for k, v in groups.items()}
    return table, roofline, fams, rows, counts


def c3_block(args, dev, hbm_peak):
    """BASELINE config 3 on one GPU: per-op device times of the PV-RCNN keypoint stage + SURVEY 8(d) per-op figures +
    the CPU composition timed beside it on a bounded sample."""
    import numpy as np
    import torch
    from vision3d_b200 import pvrcnn, synth
    B, n, N, M = 8, 100, PTS_PER_FRAME, 2048
    cfg = pvrcnn.PVRCNNConfig()
    model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0)
    stage = pvrcnn.KeypointStage(model, B, N, n, dev)
    clouds = synth.make_batch(0, B, N)
    props = pvrcnn.make_proposals(clouds, n, 0)
    grid = pvrcnn.sample_gridpoints(torch.from_numpy(props), pvrcnn.make_grid_noise(B, n, 16, 0)).reshape(B, -1, 3)
    h_pts = torch.from_numpy(np.stack(clouds)).pin_memory()
    h_grid = grid.contiguous().pin_memory()
    h_out = torch.empty((B, n, 256)).pin_memory()
    stage.load(clouds, grid)
    for _ in range(3):
        stage.step()
    torch.cuda.synchronize(dev)

    def e2e_step():
        stage.points.copy_(h_pts, non_blocking=True)
        stage.gridpoints.copy_(h_grid, non_blocking=True)
        out = stage.step()
        h_out.copy_(out, non_blocking=True)

    def timed(fn, k):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(dev)
        a.record()
        for _ in range(k):
            fn()
        b.record()
        torch.cuda.synchronize(dev)
        return a.elapsed_time(b) / k

    k = 10
    ms_dev = sorted(timed(stage.step, k) for _ in range(5))[2]
    ms_e2e = sorted(timed(e2e_step, k) for _ in range(5))[2]
    ops_t = stage.profile(iters=3)
    rows = [int(stage.eng.n_rows[lv].item()) for lv in range(4)]
    src_n = [B * N] + rows
    src_c = [1, 4, 32, 64, 64]
    table = []
    for name, us in ops_t:
        r = {"op": name, "us": round(us, 1)}
        if name == "fps+gather":   # dependent chain: point-updates/s, time per cloud, reference-style traffic for comparison
            r.update(point_updates=B * M * N, gupdates_per_s=round(B * M * N / us / 1e3, 2), us_per_cloud=round(us / B, 1),
                     hbm_min_bytes=12 * B * N + 4 * B * M, ref_style_gbs=round(20 * B * M * N / us / 1e3, 1))
        elif name.endswith("/ball_query"):
            if name.startswith("roi"):
                tests, nb = 2 * B * (n * 16) * M, 12 * B * (M + n * 16) + 4 * B * n * 16 * 48
            else:
                i = int(name[2])
                tests, nb = 2 * M * src_n[i], 12 * (src_n[i] + B * M) + 4 * B * M * 48
            r.update(tests=tests, gtests_per_s=round(tests / us / 1e3, 2), alg_bytes=nb)
        elif "fused_group" in name:
            ns = 16 if name.endswith("0") else 32
            if name.startswith("roi"):
                dims, MM = [515, 192, 96], n * 16
            else:
                i = int(name[2])
                dims, MM = [cfg.PSA_MLPS[i][0][0] + 3] + list(cfg.PSA_MLPS[i][0][1:]), M
            rows_ = B * MM * ns
            fl = 2 * rows_ * (dims[0] * dims[1] + dims[1] * dims[2])
            r.update(rows=rows_, mlp=dims, gflop=round(fl / 1e9, 2), tflops=round(fl / us / 1e6, 2),
                     grouped_bytes_not_materialised=4 * rows_ * dims[0])
        elif name == "bev_gather":
            C_bev = stage.eng.bev_nhwc.shape[1]
            nb = B * M * (4 * C_bev * 4 + 12) + 4 * B * M * C_bev    # 4 corner rows + keypoint in, C floats out
            r.update(alg_bytes=nb, gbs=round(nb / us / 1e3, 1), frac=round(nb / us / 1e3 / hbm_peak, 4))
        elif "/group_r" in name:
            ns = 16 if name.endswith("0") else 32
            if name.startswith("roi"):
                C, MM = 512, n * 16
            else:
                C, MM = src_c[int(name[2])], M
            nb = B * MM * ns * (4 * (C + 3) + 4) + 4 * B * (C + 3) * MM * ns
            r.update(alg_bytes=nb, gbs=round(nb / us / 1e3, 1), frac=round(nb / us / 1e3 / hbm_peak, 4))
        table.append(r)
    # CPU composition beside it: 1 cloud, n proposals, timed pieces (oracle C++ on 1 core, MLPs in torch on all cores)
    cpu = None
    if not args.no_cpu_baseline:
        from oracle import pvrcnn_cpu
        cpu_model = pvrcnn.init_for_benchmark(pvrcnn.PVRCNNB200(cfg), 0).eval()
        tm = {}
        t = time.perf_counter()
        pvrcnn_cpu.keypoint_stage(cpu_model, clouds[:1], grid[:1].numpy(), timings=tm)
        dt = time.perf_counter() - t
        cpu = {"value": round(1.0 / dt, 4), "unit": "frames/s", "cores": torch.get_num_threads(), "kind": "port",
               "sample": "1 cloud of the same batch through oracle/pvrcnn_cpu.py (FPS / ball query / grouping = scalar "
                         "C++ restatement on 1 core; sparse backbone, shared MLPs, grid_sample in torch on %d threads)"
                         % torch.get_num_threads(),
               "fps_s": round(tm["fps_s"], 4), "vsa_s": round(tm["vsa_s"], 3), "roi_s": round(tm["roi_s"], 3),
               "whole_stage_s": round(dt, 3)}
    total_us = sum(us for _, us in ops_t)
    keypoint_only = sum(us for nme, us in ops_t if not nme.startswith("backbone/"))
    return {"workload": "c3: PV-RCNN keypoint stage, batch 8 x %d pts: FPS-2048 + gather, sparse backbone (all levels kept), "
                        "5-source multi-radius ball query / grouping / shared MLP / max, BEV gather, RoI-grid pool (100 "
                        "proposals x 16 injected grid points), reduction MLP" % N,
            "metric": "frames/sec (PV-RCNN keypoint stage)", "value": round(B / (ms_dev / 1e3), 1), "unit": "frames/s",
            "ms_per_step": round(ms_dev, 3), "frames_per_step": B, "cuda_graph": False,
            "e2e": {"value": round(B / (ms_e2e / 1e3), 1), "unit": "frames/s", "ms_per_step": round(ms_e2e, 3),
                    "h2d_bytes_per_step": h_pts.numel() * 4 + h_grid.numel() * 4, "d2h_bytes_per_step": h_out.numel() * 4},
            "ms_keypoint_ops_only": round(keypoint_only / 1e3, 3), "ms_sum_of_ops": round(total_us / 1e3, 3),
            "mlp": "shared MLPs fused with the grouping and the max (v3d_sa_fused: tcgen05 bf16x3, fp32 accumulate, "
                   "parity-tested <= 1e-4); BEV gather = v3d_bev_gather; the 3072->256->256 reduction MLP = two torch "
                   "Linear layers (cuBLAS fp32, TF32 off)",
            "active_sites_per_level": rows, "per_op": table, "cpu_baseline": cpu}


def _conv_variant():
    """fetch scheme / L1 bypass / wait flavour of the tcgen05 kernel in this process (csrc/sparse_conv_tc.cu)"""
    from vision3d_b200 import _lib
    v = _lib.load().v3d_sparse_conv_tc_variant()
    return {"fetch_scheme": v & 7, "cp_async_cg": (v >> 3) & 1, "spin_wait": (v >> 4) & 1,
            "note": "scheme 4 (Cin = 64 layers): absent neighbours are not copied, lane-masked MMAs; other layers scheme 1"}


def run_b200(args):
    import torch
    import torch.distributed as dist
    from vision3d_b200 import _lib, second
    from vision3d_b200 import dist as vdist

    rank, world, local = vdist.init()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device: there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    lib = _lib.load()
    assert lib.v3d_check_device() == 0, "not an sm_100 device"
    h = Harness(dev, world, args.steps, args.repeats)
    B = args.batch

    # ---------------- headline: t16, weak scaling
    eng, step_device, step_e2e, h_sets, _ = second_workload(args, h, rank, world, dev, "t16", B, 100000 * rank)
    sampler = ClockSampler(local) if rank == 0 else None
    ms_dev, dev_all = h.measure(step_device)
    ms_e2e, e2e_all = h.measure(step_e2e)
    clocks = sampler.stop() if sampler else None
    frames = world * B * args.steps
    value = frames / (ms_dev / 1e3)
    e2e = frames / (ms_e2e / 1e3)

    hbm_peak, tf_peak, peak_kind = peaks()
    table = roofline = fams = rows = counts = None
    if rank == 0:
        table, roofline, fams, rows, counts = per_op_table(eng, h_sets, hbm_peak, peak_kind)
    launches_per_step = eng.kernel_launches
    dense_mb = eng.bev_nhwc.numel() * 4 / 1e6 if hasattr(eng, "bev_nhwc") else eng.dense_out.numel() * 4 / 1e6
    d2h = eng.d2h_bytes() * world
    h2d = eng.h2d_bytes()
    graph_on = eng.graph is not None
    del eng, step_device, step_e2e
    torch.cuda.empty_cache()

    # ---------------- c5: 3-class, global batch 64 sharded over the ranks (strong scaling)
    c5 = None
    if not args.no_c5:
        lo, hi = vdist.frame_range(C5_GLOBAL_BATCH, rank, world)
        fl = hi - lo
        assert fl * world == C5_GLOBAL_BATCH, "global batch 64 must divide over the ranks"
        e5, sd5, se5, hs5, _ = second_workload(args, h, rank, world, dev, "c5", fl, 7000000 + 1000003 * lo)
        ms5, all5 = h.measure(sd5, repeats=7)
        ms5e, _ = h.measure(se5, repeats=7)
        if rank == 0:
            t5 = e5.profile_ops(iters=2)
            fam5 = {}
            for name, us in t5:
                k = ("sparse_conv" if name.startswith(("subm_L", "sconv_L", "pack_L")) else
                     "rule_books" if name.startswith(("rulebook", "site_table")) else
                     "rpn(cudnn)" if name.startswith("rpn") else "head+nms+pack" if name.startswith(
                         ("heads", "nms", "pack_result")) else name.split("+")[0])
                fam5[k] = round(fam5.get(k, 0.0) + us, 1)
            c5 = {"workload": "c5: SECOND 3-class (core/config.py defaults), GLOBAL batch %d, rank r takes frames "
                              "[%d r, %d (r+1)), one in-graph all-gather of the padded detections"
                              % (C5_GLOBAL_BATCH, fl, fl),
                  "metric": "frames/sec (SECOND 3-class, global batch 64)", "scaling": "strong", "n_gpus": world,
                  "frames_per_rank": fl, "global_batch": C5_GLOBAL_BATCH, "steps": args.steps, "repeats": len(all5),
                  "value": round(C5_GLOBAL_BATCH * args.steps / (ms5 / 1e3), 2), "unit": "frames/s",
                  "ms_per_step": round(ms5 / args.steps, 4),
                  "e2e": {"value": round(C5_GLOBAL_BATCH * args.steps / (ms5e / 1e3), 2), "unit": "frames/s",
                          "ms_per_step": round(ms5e / args.steps, 4), "h2d_bytes_per_step": e5.h2d_bytes(),
                          "d2h_bytes_per_step": e5.d2h_bytes() * world},
                  "op_family_us": fam5, "active_sites_per_level": [int(e5.n_rows[lv].item()) for lv in range(5)],
                  "cuda_graph": e5.graph is not None}
        del e5, sd5, se5
        torch.cuda.empty_cache()

    # ---------------- c3: PV-RCNN keypoint stage (N = 1 only: it is a per-GPU op benchmark)
    c3 = None
    if world == 1 and not args.no_c3:
        try:
            c3 = c3_block(args, dev, hbm_peak)
        except Exception as e:  # the headline must survive a failure of a secondary block; it is reported, not hidden
            c3 = {"workload": "c3", "error": repr(e)[:500]}
        torch.cuda.empty_cache()

    line = None
    if rank == 0:
        vox = [r for r in table if r["op"].startswith("voxelize")][0]
        dn = [r for r in table if r["op"].startswith("dense")][0]
        extra_roof = {"voxelize+scatter": {"gbs": vox["gbs"], "frac": round(vox["gbs"] / hbm_peak, 4)},
                      "dense": {"gbs": dn["gbs"], "frac": round(dn["gbs"] / hbm_peak, 4)}, "families": fams}
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            fps, info = cpu_frames_per_s(3, warm=1)
            cpu = {"value": round(fps, 4), "unit": "frames/s", "cores": info["cores"], "kind": info["kind"],
                   "sample": CPU_SAMPLE % (3, info["cores"])}
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": "frames/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": round(ms_dev / args.steps, 4), "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "repeats": len(dev_all), "timing": {"stat": "median of `repeats` K-step regions",
                                                "region_ms_min_med_max": [round(dev_all[0], 3), round(ms_dev, 3),
                                                                          round(dev_all[-1], 3)],
                                                "e2e_region_ms_min_med_max": [round(e2e_all[0], 3), round(ms_e2e, 3),
                                                                              round(e2e_all[-1], 3)],
                                                "timed_s_total": round((sum(dev_all) + sum(e2e_all)) / 1e3, 2)},
            "config": {"workload": "t16: SECOND car-only (configs/second/car.yaml), KITTI-shape synthetic clouds "
                                   "(vision3d_b200.synth.make_cloud), %d pts/frame, batch %d per GPU, random "
                                   "He-init weights (no checkpoint ships with the reference)" % (PTS_PER_FRAME, B),
                       "global_batch": world * B, "parallelism": "frame-sharded dp%d" % world,
                       "l2": "no explicit flush: per-step working set (dense BEV %.0f MB + RPN activations) is "
                             "far larger than the 126 MB L2; 4 distinct input batches rotate" % dense_mb,
                       "cuda_graph": graph_on, "allgather_in_graph": bool(world > 1 and graph_on),
                       "engine_flags": {k: (v if not args.no_graph or k != "use_graph" else False)
                                        for k, v in second.BENCH_ENGINE_FLAGS.items()},
                       "sparse_conv": "exact-fp32 SIMT" if args.simt else
                       "tcgen05 kind::f16, bf16x3 split (h1*g1 + h1*g2 + h2*g1), fp32 accumulate in TMEM, all 14 layers "
                       "(the 4-channel input layer is zero-padded to 16 channels)",
                       "sparse_conv_variant": _conv_variant(),
                       "rpn": "torch/cuDNN fp32 storage, TF32 tensor-core math allowed=%s, mode=%s; parity-tested in "
                              "this mode (tests/test_gpu_bench_config.py: RPN <= 1e-2 of map scale vs fp32 CPU)"
                              % (torch.backends.cudnn.allow_tf32, args.rpn),
                       "e2e_excludes": "packing the numpy clouds into the pinned host buffer (SecondEngine.load_host)",
                       "active_sites_per_level": rows, "detections": int(counts["kept"])},
            "e2e": {"value": round(e2e, 2), "unit": "frames/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 4)},
            "gpu_launches": launches_per_step * args.steps, "gpu_launches_per_step": launches_per_step,
            "clocks": clocks, "roofline": roofline, "roofline_targets": extra_roof, "cpu_baseline": cpu,
            "workloads": {"c5": c5, "c3": c3}, "per_op": table,
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        # every rank has issued its last collective; leave without ncclCommDestroy: tearing the communicator down
        # while CUDA graphs that captured its all-gather are still alive blocked for minutes on the GPU box
        dist.barrier()
        torch.cuda.synchronize(dev)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return line


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.gpus > 1 and "WORLD_SIZE" not in os.environ:
        # convenience: re-launch under torchrun, one rank per GPU
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        raise SystemExit(subprocess.call(cmd))
    run_b200(args)


if __name__ == "__main__":
    main()
