#!/usr/bin/env bash
# Second (last) time-boxed GPU pass: the sparse-conv variant matrix (fetch scheme x L1 bypass x wait flavour) timed on
# one box, the fastest variant that passes parity is adopted for everything after it (bench, ncu, full gpu suite).
#   usage: gpurun --timeout 400 -- 'bash scripts/gpu_final2.sh r02x'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader > $OUT/${TAG}_gpu.txt 2>&1

# ---- 1. timing of every variant (scripts/conv_variant.py: per-layer CUDA-event times, eager t16 pass)
: > $OUT/${TAG}_conv_variants.txt
for combo in "1 0 0" "4 0 0" "4 1 0" "4 0 1" "4 1 1" "5 0 0" "5 1 0" "5 0 1" "1 1 0" "1 0 1" "1 1 1" "2 0 1" "2 1 1" "1 0 0"; do
  set -- $combo
  line=$(V3D_TC_FETCH=$1 V3D_TC_CG=$2 V3D_TC_WAIT=$3 timeout 60 python scripts/conv_variant.py 2>&1 | grep -h "conv total" | cut -c1-600)
  echo "fetch=$1 cg=$2 wait=$3 :: $line" >> $OUT/${TAG}_conv_variants.txt
  el "fetch=$1 cg=$2 wait=$3 :: $(echo "$line" | cut -c1-60)"
done

# ---- 2. rank, then adopt the fastest variant that passes the parity subset (baseline 1/0/0 needs no test here)
python - > $OUT/${TAG}_ranking.txt <<PY
import re
rows = []
for ln in open("$OUT/${TAG}_conv_variants.txt"):
    m = re.match(r"fetch=(\d) cg=(\d) wait=(\d) :: .*conv total (\d+) us", ln)
    if m:
        rows.append((int(m.group(4)), m.group(1), m.group(2), m.group(3)))
base = min([r[0] for r in rows if r[1:] == ("1", "0", "0")] or [10 ** 9])
seen = set()
for us, f, c, w in sorted(rows):
    if (f, c, w) in seen or (f, c, w) == ("1", "0", "0"):
        continue
    seen.add((f, c, w))
    if us < 0.99 * base:          # at least 1 % faster than the shipped kernel on the same box
        print(f, c, w, us)
print("1 0 0", base)
PY
cat $OUT/${TAG}_ranking.txt
CH_F=1; CH_C=0; CH_W=0
tries=0
while read f c w us; do
  if [[ "$f $c $w" == "1 0 0" ]]; then break; fi
  tries=$((tries + 1))
  if [[ $tries -gt 2 ]]; then break; fi
  V3D_TC_FETCH=$f V3D_TC_CG=$c V3D_TC_WAIT=$w timeout 150 python -m pytest tests/test_gpu_parity.py tests/test_gpu_bench_config.py \
      -m gpu -q -rf -p no:cacheprovider --timeout 120 \
      -k "(sparse_conv or stage_by_stage or second_stream) and not subprocess" > $OUT/${TAG}_tests_${f}${c}${w}.log 2>&1
  rc=$?
  el "parity of fetch=$f cg=$c wait=$w: rc=$rc $(tail -1 $OUT/${TAG}_tests_${f}${c}${w}.log)"
  grep -E "^FAILED|^ERROR|^E  " $OUT/${TAG}_tests_${f}${c}${w}.log | head -8
  if [[ $rc -eq 0 ]]; then CH_F=$f; CH_C=$c; CH_W=$w; break; fi
done < $OUT/${TAG}_ranking.txt
echo "$CH_F $CH_C $CH_W" > $OUT/${TAG}_chosen_variant.txt
el "chosen variant: fetch=$CH_F cg=$CH_C wait=$CH_W"
export V3D_TC_FETCH=$CH_F V3D_TC_CG=$CH_C V3D_TC_WAIT=$CH_W

# ---- 3. bench (b200 arm, then the CPU arm)
timeout 200 python bench.py > $OUT/${TAG}_bench_N1.json 2> $OUT/${TAG}_bench.err
el "bench rc=$?"
tail -2 $OUT/${TAG}_bench.err
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_N1.json").read().strip().splitlines()[-1])
    print("bench:", d["value"], d["e2e"]["value"], d["ms_per_step"], d["roofline"]["frac"], d["clocks"])
    print([(r["op"][:18], round(r["us"])) for r in d["per_op"] if r["op"].startswith(("subm", "sconv"))])
    w = d["workloads"]
    print("c5:", w["c5"] and (w["c5"]["value"], w["c5"]["ms_per_step"]))
    c3 = w["c3"]
    print("c3:", c3 and (c3.get("error") or (c3["value"], c3["ms_per_step"])))
except Exception as e:
    print("bench parse failed", e)
PY

# ---- 4. ncu evidence for the chosen variant
timeout 100 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_launches.csv \
    python scripts/ncu_step.py --steps 2 --rpn fused_nhwc > $OUT/${TAG}_ncu_list.log 2>&1
python scripts/summarize_ncu.py launches $OUT/${TAG}_launches.csv > $OUT/${TAG}_launches.md 2>&1
el "launch list done"
timeout 200 ncu --set full --import-source on --clock-control none -k 'regex:.*(sparse_conv_tc).*' \
    --launch-skip 14 --launch-count 10 -o $OUT/${TAG}_prof_conv -f python scripts/ncu_step.py --steps 2 --rpn fused_nhwc \
    > $OUT/${TAG}_ncu_conv.log 2>&1
python scripts/summarize_ncu.py counters $OUT/${TAG}_prof_conv.ncu-rep sparse_conv_tc > $OUT/${TAG}_conv_counters.md 2>&1
python scripts/summarize_ncu.py traffic $OUT/${TAG}_prof_conv.ncu-rep "profiles/${TAG}_conv_counters.md (ncu --set full, one eager step, batch 16)" > $OUT/${TAG}_traffic_conv.json 2>&1
rm -f $OUT/*.ncu-rep
el "conv counters done"
head -14 $OUT/${TAG}_conv_counters.md | cut -c1-250

# ---- 5. the whole gpu suite under the chosen variant
timeout 300 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --timeout 200 \
    --deselect tests/test_gpu_parity.py::test_voxelize_cluster_dsmem_variant_subprocess \
    > $OUT/${TAG}_tests_all.log 2>&1
el "full suite rc=$?: $(tail -1 $OUT/${TAG}_tests_all.log)"
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests_all.log | head -20

timeout 120 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/${TAG}_bench_reference.json 2>> $OUT/${TAG}_bench.err
el "reference arm: $(cut -c1-200 $OUT/${TAG}_bench_reference.json)"
du -sh $OUT
