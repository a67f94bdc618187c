#!/bin/bash
# the other BASELINE configs (C3 vector, C4 serial line, C5 haulage) at reduced event counts: relative numbers
mkdir -p gpurun_out
for wl in ${WLS:-vector_flow serial_line haulage}; do
  for lg in off ring; do
    ev=${EV:-2000}; reps=${REPS:-65536}
    timeout 600 python bench.py --workload $wl --steps 1 --warmup 1 --reps $reps --events $ev --no-cpu-baseline --e2e-steps 1 --log $lg ${LIB:+--lib $LIB} > gpurun_out/wl_${wl}_${lg}.json 2> gpurun_out/wl_${wl}_${lg}.err
    echo "== $wl log=$lg $(python -c "import sys,json; d=json.loads(open('gpurun_out/wl_${wl}_${lg}.json').readline()); print('value %.3e kernel_ms %.2f state_B %d faulted %d cfg %s' % (d['value'], d['roofline']['kernel_ms'], d['roofline']['state_bytes_per_replication'], d['faulted_replications'], d['config']['state']))" 2>&1 | tail -1)"
  done
done
