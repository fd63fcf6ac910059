#!/bin/bash
# One GPU visit of round 2:  gpurun --timeout 1500 -- 'bash tools/gpu_visit.sh TAG [legs]'
# legs (default: tests ref bench ncu): tests ref bench bench1 benchref fit ncu ceil san
#   tests  the whole parity suite (pytest -m gpu) + smoke
#   ref    tools/ref_probe.py: which reference build simulates config 2 correctly, and how long it takes
#   bench  bench.py (ours) and bench.py --impl reference
#   ncu    launch list of the bench command + `ncu --set full` captures of configs 2..5 and of the RNG ceiling kernel
# Everything lands in gpurun_out/; summaries worth keeping are copied to profiles/ by hand afterwards.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
TAG=${1:-r2a}
LEGS=${2:-tests ref bench ncu}
T0=$(date +%s)
el() { echo "== $1 rc=$2 t=$(( $(date +%s)-T0 ))s"; }
export PROCELL_WATCHDOG_S=${PROCELL_WATCHDOG_S:-120}
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv,noheader | head -8
nproc

for leg in $LEGS; do
case $leg in
tests)
  timeout 900 python -m pytest tests -m gpu -q --timeout 180 -p no:cacheprovider > gpurun_out/pytest_gpu_$TAG.log 2>&1; el pytest $?
  tail -15 gpurun_out/pytest_gpu_$TAG.log
  timeout 200 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2; el smoke $?
  ;;
ref)
  timeout 700 python tools/ref_probe.py ${REF_SIZES:-1e5,1e6} > gpurun_out/ref_probe_$TAG.log 2>&1; el ref_probe $?
  cut -c1-600 gpurun_out/ref_probe_$TAG.log | tail -12
  cp gpurun_out/ref_probe.json gpurun_out/ref_probe_$TAG.json 2>/dev/null
  ;;
bench)
  timeout 400 python bench.py > gpurun_out/bench_$TAG.json 2> gpurun_out/bench_$TAG.err; el bench $?
  cat gpurun_out/bench_$TAG.json; tail -5 gpurun_out/bench_$TAG.err
  ;;
bench1)   # the one-engine, one-stream figure beside the default (two resident engines alternating)
  timeout 200 python bench.py --engines 1 --no-cpu-baseline --no-per-config --e2e-steps 2 > gpurun_out/bench_1engine_$TAG.json 2> gpurun_out/bench_1engine_$TAG.err; el bench_1engine $?
  cut -c1-400 gpurun_out/bench_1engine_$TAG.json
  ;;
benchref)
  timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$TAG.json 2> gpurun_out/bench_ref_$TAG.err; el bench_ref $?
  cat gpurun_out/bench_ref_$TAG.json; tail -5 gpurun_out/bench_ref_$TAG.err
  ;;
fit)
  timeout 200 python tools/fitness_timing.py > gpurun_out/fitness_timing_$TAG.log 2>&1; el fitness_timing $?
  cat gpurun_out/fitness_timing_$TAG.log
  ;;
ceil)
  timeout 200 ncu --set full --clock-control none -k regex:k_rng_ceiling -c 6 -f -o gpurun_out/prof_ceilings_$TAG \
      python -c "import sys; sys.path.insert(0,'.'); from cuda_pro_cell_b200 import api; print(api.rng_ceiling_variants(0, 4096))" > gpurun_out/ncu_ceilings_$TAG.log 2>&1; el ncu_ceilings $?
  tail -2 gpurun_out/ncu_ceilings_$TAG.log
  ;;
san)
  bash tools/gpu_sanitize.sh $TAG; el sanitize $?
  ;;
ncu)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --batch 4 --no-cpu-baseline --no-per-config --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1; el ncu_launches $?
  for c in 2 3 5 4; do
    timeout 400 ncu --set full --clock-control none -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_config${c}_$TAG \
        python tools/prof_one.py $c 1.0 > gpurun_out/ncu_config${c}_$TAG.log 2>&1; el ncu_config$c $?
  done
  timeout 200 ncu --set full --clock-control none -k regex:k_rng_ceiling -s 1 -c 1 -f -o gpurun_out/prof_ceiling_$TAG \
      python -c "import sys; sys.path.insert(0,'.'); from cuda_pro_cell_b200 import api; print(api.rng_ceiling(0, 4096))" > gpurun_out/ncu_ceiling_$TAG.log 2>&1; el ncu_ceiling $?
  ls -la gpurun_out/*.ncu-rep
  ;;
esac
done
el done 0
