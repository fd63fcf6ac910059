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
bench1)   # (retired: the device-timed leg has one engine again)
  timeout 200 python bench.py --no-cpu-baseline --no-per-config --e2e-steps 2 > gpurun_out/bench_1engine_$TAG.json 2> gpurun_out/bench_1engine_$TAG.err; el bench_1engine $?
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
ab)       # A/B of the in-tree library builds named in AB_LIBS (make -C cuda_pro_cell_b200/csrc variant NAME=.. DEFS=..)
  : > gpurun_out/ab_$TAG.jsonl
  for lib in ${AB_LIBS:-libprocell_b200.so}; do
    PROCELL_LIB=$lib timeout 200 python tools/ab_knobs.py ${AB_REPS:-5} ${AB_KNOBS:-default} >> gpurun_out/ab_$TAG.jsonl 2>> gpurun_out/ab_$TAG.err; el ab_$lib $?
  done
  python - <<PY
import json
for l in open("gpurun_out/ab_$TAG.jsonl"):
    r = json.loads(l)
    print("%-34s cfg %d x%-4g %9.4f ms  %6.1f Gdiv/s  tail %7.1f us  idle/warp %6.1f us  donations %6d  crc %d" % (
        r["lib"] + "".join(" %s=%s" % kv for kv in r["knob"].items()), r["config"], r["scale"], r["ms_min"], r["Gdiv_s"], r["span_us"] - r["seed_phase_us"], r["idle_us_per_warp"], r["donations"], r["crc"]))
PY
  ;;
multi)    # needs a box with NGPU GPUs (gpurun --gpus N): the multi-GPU parity tests and BASELINE's multi-GPU configs
  NG=$(nvidia-smi -L | wc -l)
  timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q --timeout 180 -p no:cacheprovider -k "multi_gpu" > gpurun_out/pytest_multi_${NG}gpu_$TAG.log 2>&1; el pytest_multi $?
  tail -4 gpurun_out/pytest_multi_${NG}gpu_$TAG.log
  timeout 400 python tools/multi_gpu_configs.py $NG > gpurun_out/multi_gpu_configs_${NG}gpu_$TAG.log 2>&1; el multi_gpu_configs $?
  cat gpurun_out/multi_gpu_configs_${NG}gpu_$TAG.log
  ;;
scale)    # bench.py under torchrun exactly as the driver launches it, for every N in SCALE_NS (default: all GPUs of the box)
  NG=$(nvidia-smi -L | wc -l)
  for N in ${SCALE_NS:-$NG}; do
    if [ "$N" -eq 1 ]; then
      timeout 300 python bench.py --gpus 1 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1_$TAG.json 2> gpurun_out/scale_n1_$TAG.err
    else
      timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29500+N)) \
          bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/scale_n${N}_$TAG.json 2> gpurun_out/scale_n${N}_$TAG.err
    fi
    el scale_n$N $?
    tail -1 gpurun_out/scale_n${N}_$TAG.json | cut -c1-600; tail -3 gpurun_out/scale_n${N}_$TAG.err
  done
  ;;
ncu)
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$TAG.csv \
      python bench.py --steps 2 --warmup 3 --batch 4 --no-cpu-baseline --no-per-config --e2e-steps 1 > gpurun_out/ncu_launches_$TAG.log 2>&1; el ncu_launches $?
  for c in ${NCU_CONFIGS:-2 3 5 4}; do
    timeout 400 ncu --set full --clock-control none -k regex:k_proliferate_coop -s 1 -c 1 -f -o gpurun_out/prof_config${c}_$TAG \
        python tools/prof_one.py $c 1.0 > gpurun_out/ncu_config${c}_$TAG.log 2>&1; el ncu_config$c $?
  done
  ls -la gpurun_out/*.ncu-rep
  ;;
esac
done
el done 0
