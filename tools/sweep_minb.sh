#!/bin/bash
# Build variants with different register targets and time them on the GPU box (run under gpurun).
set -e
cd "$(dirname "$0")/../lte-gnu-radio-code_b200/csrc"
for cfg in "$@"; do
  IFS=, read -r small m1024 m2048 m4096 <<< "$cfg"
  make clean > /dev/null
  make -j16 EXTRA="-DOFDM_MINB_SMALL=$small -DOFDM_MINB_1024=$m1024 -DOFDM_MINB_2048=$m2048 -DOFDM_MINB_4096=$m4096" > /dev/null 2>&1
  echo "=== minblocks small=$small n1024=$m1024 n2048=$m2048 n4096=$m4096"
  (cd ../.. && python tools/quick_bench.py all 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-12s rx %.3f ms %6.0f GB/s | tx %.3f ms %6.0f GB/s' % (d['name'], d['rx_ms'], d['rx_GBs'], d['tx_ms'], d['tx_GBs']))
")
done
