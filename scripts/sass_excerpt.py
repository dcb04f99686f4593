"""profiles/sass_excerpt.txt: Blackwell-specific / asynchronous instructions per kernel of the shipped library
(cuobjdump -sass, CPU only).   python scripts/sass_excerpt.py > profiles/sass_excerpt.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "vision3d_b200", "libv3d_b200.so")
WATCH = ["UTCHMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS", "REDUX", "UCGABAR_ARV",
         "UCGABAR_WAIT", "ATOM", "ATOMS", "ATOMG", "RED", "MEMBAR", "HMMA", "BAR"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    res = []
    for n in out:
        n = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", n)
        n = re.sub(r"^void ", "", n)
        m = re.match(r"([\w:]+(?:<[^()]*?>)?)\(", n)
        res.append(m.group(1) if m else n[:90])
    return res


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = kernels.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            for w in WATCH:
                if op == w or op.startswith(w + "."):
                    cur[w] += 1
    names = demangle(list(kernels))
    print("# cuobjdump -sass vision3d_b200/libv3d_b200.so: Blackwell-specific / asynchronous instructions per kernel (static")
    print("# instruction counts in the shipped sm_100a code). UTCHMMA = tcgen05.mma, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit,")
    print("# UBLKCP = cp.async.bulk (1-D TMA), LDGSTS = cp.async, SYNCS = mbarrier ops, REDUX = warp reduce, UCGABAR = cluster")
    print("# barrier. No HMMA (legacy mma.sync) anywhere. Kernels of one template family are merged (count x instances).\n")
    merged = collections.OrderedDict()
    for name, c in zip(names, kernels.values()):
        fam = re.sub(r"<.*", "", name)
        key = (fam, tuple(sorted((k, v) for k, v in c.items() if k != "_total")))
        merged.setdefault(key, []).append((name, c["_total"]))
    for (fam, ops), inst in merged.items():
        if not ops:
            continue
        tpl = ", ".join(sorted({re.sub(r"^[^<]*", "", n) or "-" for n, _ in inst}))
        print("%-34s x%-3d %s" % (fam, len(inst), "  ".join("%s=%d" % kv for kv in ops)))
        if len(tpl) < 400 and tpl != "-":
            print("%-34s      instances: %s" % ("", tpl))
    hm = sum(c["HMMA"] for c in kernels.values())
    print("\nkernels: %d, total SASS instructions: %d, HMMA instructions: %d" % (
        len(kernels), sum(c["_total"] for c in kernels.values()), hm))


if __name__ == "__main__":
    sys.exit(main())
