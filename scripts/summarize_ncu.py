"""Turn ncu outputs into the small text summaries committed under profiles/.

  python scripts/summarize_ncu.py launches gpurun_out/r01_launches.csv > profiles/r01_launches.md
      (csv from: ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file ...)
  python scripts/summarize_ncu.py full gpurun_out/r01_prof_v3d.ncu-rep > profiles/r01_ncu_full.md
      (report from: ncu --set full --clock-control none --import-source on -o ...)
"""
import collections
import csv
import io
import re
import subprocess
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
    name = re.sub(r"^void ", "", name)
    m = re.match(r"([\w:]+(?:<[^()]*?>)?)\(", name)
    return (m.group(1) if m else name)[:80]


def launches(path, first_kernel="vox_insert"):
    txt = open(path).read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    names = [short(r["Kernel Name"]) for r in rows]
    starts = [i for i, n in enumerate(names) if first_kernel in n]
    a, b = (starts[0], starts[1]) if len(starts) > 1 else (0, len(rows))
    step = rows[a:b]
    tot = sum(float(r["Metric Value"]) for r in step) / 1e3
    print("# ncu launch list, one step (first kernel: %s), gpu__time_duration.sum per launch" % first_kernel)
    print("# cold-cache, serialised launches: compare SHARES, not absolutes. step total = %.1f us, %d launches\n"
          % (tot, len(step)))
    agg = collections.OrderedDict()
    for r in step:
        k = short(r["Kernel Name"])
        d = agg.setdefault(k, [0, 0.0])
        d[0] += 1
        d[1] += float(r["Metric Value"]) / 1e3
    print("| kernel | launches | total us | share |\n|---|---:|---:|---:|")
    for k, (n, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("| `%s` | %d | %.1f | %.1f%% |" % (k, n, us, 100 * us / tot))
    mine = sum(us for k, (n, us) in agg.items() if k.startswith("v3d::"))
    print("\nv3d kernels: %.1f us (%.1f%% of the step); cuDNN/cutlass + torch glue: %.1f us" % (mine, 100 * mine / tot,
                                                                                               tot - mine))


METRICS = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
           "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
           "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
           "sm__inst_executed_pipe_tensor.sum", "sm__warps_active.avg.pct_of_peak_sustained_active",
           "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
           "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "lts__t_bytes.sum",
           "smsp__cycles_active.avg", "sm__cycles_elapsed.max"]


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(out[out.index('"ID"'):])))
    print("# ncu --set full summary (per launch). traffic = dram read + write bytes.\n")
    cols = [c for c in rows[0].keys()]

    def col(metric):
        for c in cols:
            if c.split(" ")[0] == metric or c == metric:
                return c
        return None

    print("| kernel | grid | us | dram MB (r+w) | dram %pk | sm %pk | tensor-pipe %act | warps act % | regs |")
    print("|---|---|---:|---:|---:|---:|---:|---:|---:|")
    for r in rows[1:] if rows and not rows[0].get("ID", "").isdigit() else rows:
        try:
            def g(m, d=float("nan")):
                c = col(m)
                try:
                    return float(str(r[c]).replace(",", "")) if c and r[c] not in ("", None) else d
                except ValueError:
                    return d
            t = g("gpu__time_duration.sum")
            unit_row = rows[0]
            tu = unit_row.get(col("gpu__time_duration.sum"), "")
            us = t / 1e3 if "ns" in tu or t > 1e4 else t
            rb, wb = g("dram__bytes_read.sum"), g("dram__bytes_write.sum")
            bu = unit_row.get(col("dram__bytes_read.sum"), "")
            scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}.get(bu, 1e-6)
            print("| `%s` | %s | %.1f | %.2f | %.1f | %.1f | %.1f | %.1f | %d |" % (
                short(r["Kernel Name"]), r.get("Grid Size", ""), us, (rb + wb) * scale,
                g("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
                g("sm__throughput.avg.pct_of_peak_sustained_elapsed"),
                g("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
                g("sm__warps_active.avg.pct_of_peak_sustained_active"), int(g("launch__registers_per_thread", 0))))
        except Exception as e:  # keep going: one odd row must not hide the rest
            print("| %s | parse error %s |" % (r.get("Kernel Name", "?")[:40], e))


FAMILIES = [("sparse_conv_fwd", ("sparse_conv_",)), ("voxelize+vfe", ("vox_",)),
            ("dense", ("dense_",)), ("nms_rotated", ("nms_",)), ("rulebook", ("rule_", "conv_mark", "conv_scan", "conv_rank"))]


def traffic(path, tag=None):
    """profiles/traffic.json: measured DRAM bytes (read + write) per launch for each kernel family, averaged over
    the launches of the capture -- bench.py copies the dominant family's figure into roofline.traffic."""
    import json
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(out[out.index('"ID"'):])))
    units, rows = rows[0], rows[1:]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    res = {}
    for fam, keys in FAMILIES:
        tot, n, us = 0.0, 0, 0.0
        for r in rows:
            if any(k in r["Kernel Name"] for k in keys):
                for m in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
                    tot += float(r[m].replace(",", "")) * scale.get(units[m], 1.0)
                t = float(r["gpu__time_duration.sum"].replace(",", ""))
                us += t / 1e3 if units["gpu__time_duration.sum"] in ("ns", "nsecond") else t
                n += 1
        if n:
            res[fam] = {"dram_bytes_per_launch": int(tot / n), "launches": n, "avg_us_under_ncu": round(us / n, 1),
                        "source": tag or path}
    print(json.dumps(res, indent=1))


COUNTERS = [("us", "gpu__time_duration.sum"), ("cycles", "sm__cycles_elapsed.max"),
            ("tensor %act", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"),
            ("l1tex data-pipe %", "l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed"),
            ("lts %", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
            ("dram %", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
            ("L2->L1 MB", "l1tex__m_xbar2l1tex_read_bytes.sum"),
            ("LDGSTS inst", "smsp__inst_executed_op_ldgsts.sum"),
            ("smem wavefronts (lsu)", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum"),
            ("of which ldgsts", "smsp__sass_l1tex_data_pipe_lsu_wavefronts_mem_shared_op_ldgsts.sum"),
            ("of which ld", "l1tex__data_pipe_lsu_wavefronts_mem_shared_op_ld.sum"),
            ("tensor-core smem wavefronts", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum"),
            ("L2 hit %", "lts__t_sector_hit_rate.pct")]


def counters(path, pattern="sparse_conv_tc"):
    """Per-launch L1TEX / L2 / tensor counters of the kernels matching `pattern` (the evidence behind DESIGN 4.1)."""
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.DictReader(io.StringIO(out[out.index('"ID"'):])))
    units, rows = rows[0], rows[1:]
    scale = {"Gbyte": 1e3, "Mbyte": 1.0, "Kbyte": 1e-3, "byte": 1e-6}
    print("# ncu --set full counters per launch (%s), kernels matching '%s'\n" % (path.split("/")[-1], pattern))
    print("| kernel | " + " | ".join(n for n, _ in COUNTERS) + " |")
    print("|---|" + "---:|" * len(COUNTERS))
    for r in rows:
        if pattern not in r["Kernel Name"]:
            continue
        vals = []
        for n, m in COUNTERS:
            if m not in r or r[m] in ("", None):
                vals.append("n/a")
                continue
            v = float(r[m].replace(",", ""))
            if units[m] in scale:
                v *= scale[units[m]]
            if units[m] in ("ns", "nsecond"):
                v /= 1e3
            vals.append("%.1f" % v if v < 1e4 else "%.3g" % v)
        print("| `%s` | %s |" % (short(r["Kernel Name"]), " | ".join(vals)))


if __name__ == "__main__":
    {"launches": launches, "full": full, "traffic": traffic, "counters": counters}[sys.argv[1]](*sys.argv[2:])
