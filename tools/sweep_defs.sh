#!/bin/bash
# usage: sweep_defs.sh "<nvcc -D flags>" ...   builds each variant and prints quick_bench lines
set -e
cd "$(dirname "$0")/../lte-gnu-radio-code_b200/csrc"
for defs in "$@"; do
  make clean > /dev/null
  make -j16 EXTRA="$defs" > /dev/null 2>&1 || { echo "build failed: $defs"; continue; }
  echo "=== $defs"
  (cd ../.. && python tools/quick_bench.py ${BENCH_SET:-all} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-12s rx %.3f ms %6.0f GB/s | tx %.3f ms %6.0f GB/s' % (d['name'], d['rx_ms'], d['rx_GBs'], d['tx_ms'], d['tx_GBs']))
")
done
make clean > /dev/null; make -j16 > /dev/null 2>&1
