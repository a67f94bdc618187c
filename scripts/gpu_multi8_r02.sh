#!/bin/bash
# round 2, all 8 GPUs of one box (charged 8x: kept short): the multi-GPU tests (qk_multi_* behind the C ABI on every GPU),
# then the headline at N = 8 the way the driver launches it -- C2's 1,048,576 replications split over the GPUs (strong),
# with the log ring and statistics only -- and C5's 262,144.
N=${1:-8}
mkdir -p gpurun_out/multi$N
O=gpurun_out/multi$N
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 200 python -m pytest tests/test_gpu_multi.py -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multi.log
show() { python -c "import json,sys; d=json.loads(open('$1').readline()); print('$1'.split('/')[-1], 'n_gpus %d scaling %s value %.4e per_gpu %.4e ms %.1f e2e %.4e spread %.3f chunks %s' % (d['n_gpus'], d['scaling'], d['value'], d['per_gpu_events_per_s'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks']))" 2>&1 | tail -1; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 300 $T --master-port 29612 bench.py --gpus $N --no-lockstep-probe > $O/strong_c2_n$N.json 2> $O/strong_c2_n$N.err; show $O/strong_c2_n$N.json
timeout 300 $T --master-port 29613 bench.py --gpus $N --no-lockstep-probe --log off > $O/strong_c2_off_n$N.json 2> $O/strong_c2_off_n$N.err; show $O/strong_c2_off_n$N.json
timeout 300 $T --master-port 29615 bench.py --gpus $N --workload haulage --no-lockstep-probe > $O/strong_c5_n$N.json 2> $O/strong_c5_n$N.err; show $O/strong_c5_n$N.json
