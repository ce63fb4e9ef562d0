#!/bin/bash
# Run every GPU test node in its own process (a trapped kernel poisons the CUDA context of its process only),
# each under a timeout, and collect a summary in gpurun_out/battery_<tag>.txt.
# usage: tools/gpu_battery.sh <tag> [pytest -k expression] [test file]
TAG=${1:-run}
KEXPR=${2:-}
FILE=${3:-tests/test_kernels_gpu.py}
OUT=gpurun_out/battery_${TAG}.txt
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,driver_version,clocks.sm,clocks.max.sm --format=csv > "$OUT" 2>&1
if [ -n "$KEXPR" ]; then
  IDS=$(python -m pytest "$FILE" -m gpu -k "$KEXPR" --collect-only -q 2>/dev/null | grep "::")
else
  IDS=$(python -m pytest "$FILE" -m gpu --collect-only -q 2>/dev/null | grep "::")
fi
PASS=0; FAIL=0
for id in $IDS; do
  LOG=$(timeout 180 python -m pytest "$id" -m gpu -x -q -s 2>&1)
  RC=$?
  if [ $RC -eq 0 ]; then PASS=$((PASS+1)); STATUS=PASS; else FAIL=$((FAIL+1)); STATUS="FAIL($RC)"; fi
  echo "== $STATUS $id" >> "$OUT"
  echo "$LOG" | grep -E "^\[|Error|error|assert|Mismatch|mismatch|nan" | head -12 >> "$OUT"
done
echo "TOTAL pass=$PASS fail=$FAIL" >> "$OUT"
tail -1 "$OUT"
