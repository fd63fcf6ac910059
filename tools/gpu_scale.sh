#!/bin/bash
# scaling run on one 8-GPU box: single-process NCCL tests + bench at N = 1, 2, 4, 8 as the driver launches it
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | wc -l
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q --timeout 90 -k "multi_gpu" 2>&1 | tail -2
for n in 1 2 4 8; do
  if [ $n -eq 1 ]; then
    timeout 300 python bench.py --gpus 1 --steps 200 --warmup 5 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2961$n bench.py --gpus $n --steps 200 --warmup 5 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "bench n=$n rc=$?"; python -c "
import json
d=json.loads(open('gpurun_out/scale_n$n.json').read().strip().splitlines()[-1])
print('N=%d value=%.4g ms/step=%.3f e2e=%.4g frac=%.3f'%(d['n_gpus'],d['value'],d['ms_per_step'],d['e2e']['value'],d['roofline']['frac']))" || tail -5 gpurun_out/scale_n$n.err
done
