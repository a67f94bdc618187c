#!/bin/bash
# final build on 2 GPUs of one box: the multi-GPU tests (qk_multi_* behind the C ABI: tensor maps per device, the N-shard
# summary against the single-batch answer) and the headline split over the two GPUs (strong scaling), as the driver
# launches it
N=2
mkdir -p gpurun_out/multi2b
O=gpurun_out/multi2b
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 600 python -m pytest tests/test_gpu_multi.py -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multi.log
show() { python -c "import json,sys; d=json.loads(open('$1').readline()); print('$1'.split('/')[-1], 'n_gpus %d scaling %s value %.4e per_gpu %.4e ms %.1f e2e %.4e spread %.3f chunks %s' % (d['n_gpus'], d['scaling'], d['value'], d['per_gpu_events_per_s'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks']))" 2>&1 | tail -1; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 $T --master-port 29612 bench.py --gpus $N --no-lockstep-probe > $O/strong_c2_n$N.json 2> $O/strong_c2_n$N.err; show $O/strong_c2_n$N.json
timeout 600 $T --master-port 29613 bench.py --gpus $N --reps 262144 --no-lockstep-probe > $O/strong_c2_quarter_n$N.json 2> $O/strong_c2_quarter_n$N.err; show $O/strong_c2_quarter_n$N.json
