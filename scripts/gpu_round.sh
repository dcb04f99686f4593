#!/usr/bin/env bash
# One-stop GPU pass, meant to be the COMMAND of a single gpurun call (everything it writes stays well under the
# 64 MiB gpurun_out limit):   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh r02a [tests] [bench] [ncu]'
#   tests: pytest -m gpu       bench: bench.py (b200 arm + reference arm)
#   ncu  : launch list of one eager step + `--set full` capture of the sparse-conv / voxelize / rule-book kernels
# Summaries: scripts/summarize_ncu.py (run locally on the .ncu-rep / csv that come back).
set -u
TAG=${1:-rXX}
shift || true
WHAT=${*:-tests bench ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [[ " $WHAT " == *" tests "* ]]; then
  timeout 2400 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 900 -s 2>&1 | tail -60 | tee $OUT/${TAG}_tests.log
fi
if [[ " $WHAT " == *" bench "* ]]; then
  timeout 900 python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench.err
  tail -5 $OUT/${TAG}_bench.err
  timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
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
fi
if [[ " $WHAT " == *" ncu "* ]]; then
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
      python scripts/ncu_step.py --steps 2 --rpn fused_nhwc > $OUT/${TAG}_ncu_list.log 2>&1
  timeout 900 ncu --set full --import-source on --clock-control none -k 'regex:.*(sparse_conv_tc).*' \
      --launch-skip 14 --launch-count 14 -o $OUT/${TAG}_prof_conv -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
      > $OUT/${TAG}_ncu_conv.log 2>&1
  timeout 600 ncu --set full --clock-control none -k 'regex:.*(vox_|table_|rule_|conv_mark|conv_scan|conv_rank|dense_|feature_pack|rulebook).*' \
      --launch-skip 40 --launch-count 40 -o $OUT/${TAG}_prof_rest -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
      > $OUT/${TAG}_ncu_rest.log 2>&1
fi
du -sh $OUT
