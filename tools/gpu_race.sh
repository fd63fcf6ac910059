#!/bin/bash
# compute-sanitizer racecheck + initcheck on the small runs of tools/gpu_last.sh
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
sed -n '/^cat > \/tmp\/san.py/,/^PY$/p' tools/gpu_last.sh | sed '1d;$d' > /tmp/san.py
for tool in racecheck initcheck; do
  timeout 25 /usr/local/cuda/bin/compute-sanitizer --tool $tool --error-exitcode 9 python /tmp/san.py > gpurun_out/sanitizer_${tool}_last.log 2>&1; echo "$tool rc=$?"
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY|^ok" gpurun_out/sanitizer_${tool}_last.log | head -4
done
grep -E "hazard" gpurun_out/sanitizer_racecheck_last.log | sed 's/0x[0-9a-f]*//g' | sort | uniq -c | sort -rn | head -12
