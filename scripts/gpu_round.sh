#!/usr/bin/env bash
# One-stop GPU pass, meant to be the COMMAND of a single gpurun call (everything it writes stays well under the
# 64 MiB gpurun_out limit):   gpurun --timeout 1800 -- 'bash scripts/gpu_round.sh r02z [tests] [bench] [ncu]'
#   tests: pytest -m gpu       bench: bench.py (b200 arm + reference arm)
#   ncu  : launch lists of one eager SECOND step and one C3 step + `--set full` captures of the tensor-core kernels
#          (sparse conv, fused set abstraction) and section captures of the rest.
# Summaries: scripts/summarize_ncu.py (run locally on the .ncu-rep / csv that come back).
set -u
TAG=${1:-rXX}
shift || true
WHAT=${*:-tests bench ncu}
OUT=gpurun_out
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/${TAG}_gpu.txt 2>&1
if [[ " $WHAT " == *" tests "* ]]; then
  timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider --timeout 300 -s > $OUT/${TAG}_tests.log 2>&1
  grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  |err |Timeout" $OUT/${TAG}_tests.log | head -40
fi
if [[ " $WHAT " == *" bench "* ]]; then
  timeout 600 python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench.err
  tail -3 $OUT/${TAG}_bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_N1.json").read().strip().splitlines()[-1])
    print("bench:", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    print([(r["op"][:18], round(r["us"])) for r in d["per_op"]])
    w = d["workloads"]
    print("c5:", w["c5"] and (w["c5"]["value"], w["c5"]["ms_per_step"], w["c5"]["op_family_us"]))
    c3 = w["c3"]
    print("c3:", c3 and (c3.get("error") or (c3["value"], c3["ms_per_step"], c3["cpu_baseline"])))
    print("cpu:", d["cpu_baseline"])
    print("ref:", open("$OUT/${TAG}_bench_reference.json").read()[:300])
except Exception as e:
    print("bench parse failed", e)
PY
fi
if [[ " $WHAT " == *" ncu "* ]]; then
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
      python scripts/ncu_step.py --steps 2 --rpn fused_nhwc > $OUT/${TAG}_ncu_list.log 2>&1
  timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_c3_launches.csv \
      python scripts/ncu_c3.py 2 > $OUT/${TAG}_ncu_c3_list.log 2>&1
  timeout 500 ncu --set full --import-source on --clock-control none -k 'regex:.*(sparse_conv_tc).*' \
      --launch-skip 14 --launch-count 14 -o $OUT/${TAG}_prof_conv -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
      > $OUT/${TAG}_ncu_conv.log 2>&1
  timeout 400 ncu --set full --clock-control none -k 'regex:.*(vox_|table_|rule_|conv_mark|conv_scan|conv_rank|dense_|feature_pack|nms_|topk_rows|cls_logits|reg_gather|head_decode|pack_kernel).*' \
      --launch-skip 52 --launch-count 52 -o $OUT/${TAG}_prof_rest -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
      > $OUT/${TAG}_ncu_rest.log 2>&1
  timeout 500 ncu --set full --clock-control none -k 'regex:.*(sa_fused|ball_query|fps_|bq_bounds|bs_|bev_gather|query_group|sa_pack|to_global|batch_offsets).*' \
      --launch-skip 40 --launch-count 40 -o $OUT/${TAG}_prof_c3 -f python scripts/ncu_c3.py 2 \
      > $OUT/${TAG}_ncu_c3.log 2>&1
  # summarise on the box and leave the (large) reports behind: gpurun_out is limited to 64 MiB
  python scripts/summarize_ncu.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.md 2>&1
  python scripts/summarize_ncu.py launches $OUT/${TAG}_c3_launches.csv fps_ > $OUT/${TAG}_c3_launches.md 2>&1
  python scripts/summarize_ncu.py counters $OUT/${TAG}_prof_conv.ncu-rep sparse_conv_tc > $OUT/${TAG}_conv_counters.md 2>&1
  python scripts/summarize_ncu.py counters $OUT/${TAG}_prof_c3.ncu-rep "" > $OUT/${TAG}_c3_counters.md 2>&1
  python scripts/summarize_ncu.py full $OUT/${TAG}_prof_rest.ncu-rep > $OUT/${TAG}_ncu_full.md 2>&1
  python scripts/summarize_ncu.py full $OUT/${TAG}_prof_conv.ncu-rep >> $OUT/${TAG}_ncu_full.md 2>&1
  python scripts/summarize_ncu.py traffic $OUT/${TAG}_prof_conv.ncu-rep "profiles/${TAG}_ncu_full.md (ncu --set full, one eager step, batch 16)" > $OUT/${TAG}_traffic_conv.json 2>&1
  python scripts/summarize_ncu.py traffic $OUT/${TAG}_prof_rest.ncu-rep "profiles/${TAG}_ncu_full.md (ncu --set full, one eager step, batch 16)" > $OUT/${TAG}_traffic_rest.json 2>&1
  rm -f $OUT/*.ncu-rep
fi
du -sh $OUT
