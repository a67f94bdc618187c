#!/bin/bash
# weak scaling on one box: N = 1, 2, 4, 8 back to back (the driver computes efficiency from these lines itself)
mkdir -p gpurun_out
for n in ${NS:-1 2 4 8}; do
  if [ $n -eq 1 ]; then
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  else
    python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 3 --warmup 3 > gpurun_out/scale_n$n.json 2> gpurun_out/scale_n$n.err
  fi
  echo "N=$n rc=$? $(python -c "import json; d=json.loads(open('gpurun_out/scale_n$n.json').readline()); print('value %.4e per_gpu %.4e ms %.1f e2e %.4e' % (d['value'], d['per_gpu_events_per_s'], d['ms_per_step'], d['e2e']['value']))" 2>&1 | tail -1)"
done
