#!/bin/bash
# multi-GPU visit: GPU tests (incl. the single-process NCCL path) + bench under torchrun, as the driver launches it
cd "$(dirname "$0")/.."
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 400 python -m pytest tests -m gpu -x -q --timeout 90 > gpurun_out/pytest_gpu_n$N.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu_n$N.log
for n in $(seq 1 $N); do
  if [ $n -eq 1 ] || [ $n -eq 2 ] || [ $n -eq 4 ] || [ $n -eq 8 ]; then
    if [ $n -eq 1 ]; then
      timeout 600 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    else
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
    fi
    echo "bench n=$n rc=$?"; python -c "
import json,sys
d=json.loads(open('gpurun_out/bench_n$n.json').read().strip().splitlines()[-1])
print('N=%d value=%.4g ms/step=%.3f e2e=%.4g frac=%.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']))" || tail -5 gpurun_out/bench_n$n.err
  fi
done
