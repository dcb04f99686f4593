set -u
OUT=gpurun_out; mkdir -p $OUT
timeout 400 python -m pytest tests/test_gpu_pvrcnn.py -m gpu -q -p no:cacheprovider --timeout 120 -s > $OUT/r02m_tests.log 2>&1
grep -E "passed|failed|^FAILED|^ERROR|Error|assert |^E  |err |Timeout" $OUT/r02m_tests.log | head -20
timeout 300 python bench.py --no-c5 --no-cpu-baseline --repeats 3 > $OUT/r02m_bench.json 2>$OUT/r02m_bench.err
python - <<PY
import json
d = json.loads(open("$OUT/r02m_bench.json").read().strip().splitlines()[-1])
c3 = d["workloads"]["c3"]
print("c3:", c3.get("error") or (c3["value"], c3["ms_per_step"], c3["ms_keypoint_ops_only"]))
print([(r["op"], r["us"]) for r in c3["per_op"] if "ball_query" in r["op"]])
PY
