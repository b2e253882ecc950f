#!/bin/bash
set -e
cd "$(dirname "$0")/../lte-gnu-radio-code_b200/csrc"
for cfg in "$@"; do
  IFS=, read -r small m1024 m2048 m4096 <<< "$cfg"
  make clean > /dev/null
  make -j16 EXTRA="-DOFDM_MINB_TX_SMALL=$small -DOFDM_MINB_TX_1024=$m1024 -DOFDM_MINB_TX_2048=$m2048 -DOFDM_MINB_TX_4096=$m4096" > /dev/null 2>&1
  echo "=== TX minblocks small=$small n1024=$m1024 n2048=$m2048 n4096=$m4096"
  grep -h -A1 "tx_kernelILi6ELi4E\|tx_kernelILi11ELi6E\|tx_kernelILi10ELi4E\|tx_kernelILi12ELi4E" build/*.ptxas.log | grep -E "spill" | tr '\n' ' '; echo
  (cd ../.. && python tools/quick_bench.py all 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    print('%-12s tx %.3f ms %6.0f GB/s' % (d['name'], d['tx_ms'], d['tx_GBs']))
")
done
