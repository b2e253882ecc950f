#!/bin/bash
# usage (on the GPU box): BENCH_SET=large TESTS=1 tools/sweep_variants.sh "<nvcc -D flags>" ...
# Builds each variant, optionally runs the GPU parity tests on it, and prints quick_bench lines.
cd "$(dirname "$0")/../lte-gnu-radio-code_b200/csrc"
for defs in "$@"; do
  make clean > /dev/null
  if ! make -j32 EXTRA="$defs" > /tmp/build.log 2>&1; then echo "=== build failed: $defs"; tail -5 /tmp/build.log; continue; fi
  echo "=== [$defs]  spills: $(grep -h 'spill' build/*.ptxas.log | grep -vc ' 0 bytes spill stores')"
  if [ -n "$TESTS" ]; then (cd ../.. && python -m pytest tests -m gpu -x -q 2>&1 | tail -3); fi
  (cd ../.. && python tools/quick_bench.py ${BENCH_SET:-all} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-12s rx %.3f ms %6.0f GB/s | tx %.3f ms %6.0f GB/s | ber %.2e locked %.3f' % (d['name'], d['rx_ms'], d['rx_GBs'], d['tx_ms'], d['tx_GBs'], d['ber'], d['locked']))
")
done
make clean > /dev/null; make -j32 > /dev/null 2>&1
