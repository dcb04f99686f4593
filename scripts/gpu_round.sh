#!/usr/bin/env bash
# One-stop GPU pass, meant to be the COMMAND of a single gpurun call (everything it writes stays well under the
# 64 MiB gpurun_out limit):   gpurun --timeout 1500 -- 'bash scripts/gpu_round.sh r02'
#   1. GPU tests   2. bench (b200 arm + reference arm)   3. ncu launch list   4. ncu --set full of the conv /
#   voxelize / dense kernels + section capture of the rest.  Summaries: scripts/summarize_ncu.py (run locally).
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 | tee $OUT/${TAG}_tests.log
timeout 600 python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/${TAG}_bench_N1.json").read().strip().splitlines()[-1])
print("bench:", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"])
print([(r["op"][:18], round(r["us"])) for r in d["per_op"]])
PY
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python scripts/ncu_step.py --steps 2 --rpn fused_nhwc > $OUT/${TAG}_ncu_list.log 2>&1
timeout 600 ncu --set full --clock-control none -k 'regex:.*(vox_|sparse_conv|dense_|feature_pack).*' \
    --launch-skip 22 --launch-count 22 -o $OUT/${TAG}_prof_a -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
    > $OUT/${TAG}_ncu_a.log 2>&1
timeout 600 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats --section Occupancy \
    --metrics dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k 'regex:.*(table_|rule_|conv_mark|conv_scan|conv_rank|nms_|head_decode|pack_kernel|cls_logits|topk_rows|reg_gather).*' \
    --launch-skip 35 --launch-count 35 -o $OUT/${TAG}_prof_b -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
    > $OUT/${TAG}_ncu_b.log 2>&1
du -sh $OUT
