#!/bin/bash
# round 2, N GPUs of one box (N = $1, default 2): the multi-GPU tests (qk_multi_* behind the C ABI, the N-shard summary
# against the single-batch answer), then strong scaling of C2 and C5: the job's replications split over the GPUs
N=${1:-2}
mkdir -p gpurun_out/multi$N
O=gpurun_out/multi$N
nvidia-smi --query-gpu=index,name --format=csv > $O/gpus.txt
timeout 900 python -m pytest tests/test_gpu_multi.py -q > $O/pytest_multi.log 2>&1; echo "pytest rc=$?"; tail -3 $O/pytest_multi.log
show() { python -c "import json,sys; d=json.loads(open('$1').readline()); print('$1'.split('/')[-1], 'n_gpus %d scaling %s value %.4e per_gpu %.4e ms %.1f e2e %.4e spread %.3f chunks %s' % (d['n_gpus'], d['scaling'], d['value'], d['per_gpu_events_per_s'], d['ms_per_step'], d['e2e']['value'], d['e2e']['spread'], d['placement']['step_chunks']))" 2>&1 | tail -1; }
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 python bench.py --no-cpu-baseline --no-lockstep-probe > $O/strong_c2_n1.json 2> $O/strong_c2_n1.err; show $O/strong_c2_n1.json
timeout 900 $T --master-port 29612 bench.py --gpus $N --no-lockstep-probe > $O/strong_c2_n$N.json 2> $O/strong_c2_n$N.err; show $O/strong_c2_n$N.json
timeout 900 $T --master-port 29613 bench.py --gpus $N --no-lockstep-probe --log off > $O/strong_c2_off_n$N.json 2> $O/strong_c2_off_n$N.err; show $O/strong_c2_off_n$N.json
timeout 900 $T --master-port 29614 bench.py --gpus $N --no-lockstep-probe --scaling weak > $O/weak_c2_n$N.json 2> $O/weak_c2_n$N.err; show $O/weak_c2_n$N.json
timeout 900 python bench.py --workload haulage --no-cpu-baseline --no-lockstep-probe > $O/strong_c5_n1.json 2> $O/strong_c5_n1.err; show $O/strong_c5_n1.json
timeout 900 $T --master-port 29615 bench.py --gpus $N --workload haulage --no-lockstep-probe > $O/strong_c5_n$N.json 2> $O/strong_c5_n$N.err; show $O/strong_c5_n$N.json
