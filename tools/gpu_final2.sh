#!/bin/bash
# round-end GPU visit: parity suite, smoke, bench (ours + reference arm), ncu launch list + full capture of the bench
# step, all five configs at full size, A/B against the previous build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r1g}
T0=$(date +%s)
el() { echo "$1 rc=$2 t=$(( $(date +%s)-T0 ))s"; }
timeout 200 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu_$TAG.log 2>&1; el pytest $?
tail -3 gpurun_out/pytest_gpu_$TAG.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; el smoke $?
timeout 300 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; el bench $?
cat gpurun_out/bench_$TAG.json; tail -3 gpurun_out/bench_$TAG.err
timeout 200 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; el bench_ref $?
cat gpurun_out/bench_ref_$TAG.json; tail -3 gpurun_out/bench_ref_$TAG.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_$TAG.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1; el ncu_launches $?
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_proliferate_coop -s 3 -c 1 -f -o gpurun_out/prof_bench_$TAG \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --e2e-steps 1 > gpurun_out/ncu_full_$TAG.log 2>&1; el ncu_full $?
timeout 300 python tools/run_all_configs.py > gpurun_out/all_configs_$TAG.log 2>&1; el all_configs $?
cut -c1-400 gpurun_out/all_configs_$TAG.log
if [ -f cuda_pro_cell_b200/libprocell_b200_base.so ]; then
  for lib in libprocell_b200_base.so libprocell_b200.so; do PROCELL_LIB=$lib timeout 100 python tools/ab_knobs.py 7 default; done > gpurun_out/ab_$TAG.jsonl 2> gpurun_out/ab_$TAG.err
  python - <<PY
import json
for l in open("gpurun_out/ab_$TAG.jsonl"):
    r=json.loads(l); print(r["lib"], r["config"], r["scale"], "%.4f ms"%r["ms_min"], "%.1f Gdiv/s"%r["Gdiv_s"], r["crc"])
PY
fi
el done 0
