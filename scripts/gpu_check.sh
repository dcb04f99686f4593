#!/usr/bin/env bash
# quick GPU pass: all gpu tests (full log kept, failures summarised) + bench (b200 arm)
set -u
TAG=${1:-rXX}
OUT=gpurun_out; mkdir -p $OUT
timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -s > $OUT/${TAG}_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  |err " $OUT/${TAG}_tests.log | head -60
timeout 900 python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench.err
tail -3 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_N1.json").read().strip().splitlines()[-1])
    print("bench:", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    print([(r["op"][:18], round(r["us"])) for r in d["per_op"]])
    w = d["workloads"]
    print("c5:", w["c5"] and (w["c5"]["value"], w["c5"]["ms_per_step"], w["c5"]["op_family_us"]))
    c3 = w["c3"]
    print("c3:", c3 and (c3.get("error") or (c3["value"], c3["ms_per_step"], [(r["op"], r["us"]) for r in c3["per_op"] if not r["op"].startswith("backbone")], c3["cpu_baseline"])))
except Exception as e:
    print("bench parse failed", e)
PY
