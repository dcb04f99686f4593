"""Aggregate warp-stall samples of an .ncu-rep per CUDA source line (ncu --import-source on, -lineinfo).
    python scripts/ncu_lines.py report.ncu-rep [kernel_index] [min_samples]
"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
kidx = int(sys.argv[2]) if len(sys.argv) > 2 else 0
min_s = int(sys.argv[3]) if len(sys.argv) > 3 else 40
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
allst = [i for i, r in enumerate(rows) if r and r[0] == "Function Name"]
starts = [i for i in allst if rows[i - 1][1].endswith(".cu")]  # one section per (kernel, source file)
print("kernels:", len(starts))
s = starts[kidx]
e = min([i for i in allst if i > s] + [len(rows) + 1]) - 1
print(rows[s][1][:120])
hdr = rows[s + 1]
iS = hdr.index("# Samples")
stall = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in rows[s + 2:e]:
    if r and r[0] not in ("", "File Path", "Function Name", "Line No"):
        try:
            top = sorted(((int(r[c]), hdr[c][6:]) for c in stall), reverse=True)[:2]
            agg[int(r[0])] = (int(r[iS]), r[1][:100], top)
        except ValueError:
            pass
tot = sum(v[0] for v in agg.values())
print("total samples", tot)
for ln, (n, src, top) in sorted(agg.items()):
    if n >= min_s:
        print("%4d %6d %5.1f%%  %-100s %s" % (ln, n, 100.0 * n / tot, src, top))
