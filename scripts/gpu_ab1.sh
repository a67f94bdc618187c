#!/bin/bash
# one pass over several builds of the library at full size.  LIBS="a.so b.so ..."
for lib in ${LIBS:-quokkasim_b200/libquokka_b200.so}; do
  for lg in off ring; do
    echo "== $lib log=$lg $(timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline --e2e-steps 1 --log $lg --lib $lib 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.readline()); print('value %.4e kernel_ms %.2f' % (d['value'], d['roofline']['kernel_ms']))")"
  done
done
