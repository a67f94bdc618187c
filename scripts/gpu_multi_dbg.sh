#!/bin/bash
mkdir -p gpurun_out/multidbg
O=gpurun_out/multidbg
g++ -std=c++17 -O1 -Iinclude examples/discrete_queue_multi.cpp -o /tmp/dqm -Lquokkasim_b200 -l:libquokka_b200.so -Wl,-rpath,$PWD/quokkasim_b200
for g in 1 2; do /tmp/dqm --replications 30001 --events 2000 --gpus $g > $O/out_g$g.txt 2>&1; echo "rc=$?"; done
diff $O/out_g1.txt $O/out_g2.txt
for g in 1 2; do /tmp/dqm --replications 30000 --events 2000 --gpus $g > $O/outb_g$g.txt 2>&1; done
diff $O/outb_g1.txt $O/outb_g2.txt | head
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29633 bench.py --gpus 2 --no-lockstep-probe --log=off > $O/strong_c2_off_n2.json 2> $O/strong_c2_off_n2.err; python -c "import json; d=json.loads(open('$O/strong_c2_off_n2.json').readline()); print('off n2 value %.4e e2e %.4e'%(d['value'],d['e2e']['value']))"
