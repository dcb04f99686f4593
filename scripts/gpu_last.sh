#!/usr/bin/env bash
# Last short GPU pass (defaults, no V3D_TC_* environment): whole gpu suite, smoke(), ncu launch list of one C3 step.
#   usage: gpurun --timeout 95 -- 'bash scripts/gpu_last.sh r02w'
set -u
TAG=${1:-rXX}
OUT=gpurun_out
mkdir -p $OUT
T0=$(date +%s)
el() { echo "[$(( $(date +%s) - T0 )) s] $*"; }
timeout 70 python -m pytest tests -m gpu -q -rf -p no:cacheprovider --timeout 60 \
    --deselect tests/test_gpu_parity.py::test_voxelize_cluster_dsmem_variant_subprocess > $OUT/${TAG}_tests_all.log 2>&1
el "full suite (defaults) rc=$?: $(tail -1 $OUT/${TAG}_tests_all.log)"
grep -E "^FAILED|^ERROR" $OUT/${TAG}_tests_all.log | head -10
timeout 40 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $OUT/${TAG}_c3_launches.csv \
    python scripts/ncu_c3.py 2 > $OUT/${TAG}_ncu_c3_list.log 2>&1
python scripts/summarize_ncu.py launches $OUT/${TAG}_c3_launches.csv fps_ > $OUT/${TAG}_c3_launches.md 2>&1
el "c3 launch list rc=$?"
head -12 $OUT/${TAG}_c3_launches.md
timeout 40 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $OUT/${TAG}_smoke.log 2>&1
el "smoke: $(tail -1 $OUT/${TAG}_smoke.log)"
