#!/bin/bash
# usage (on the GPU box): TESTS="pad padlut" BENCH_SET=all tools/run_variants.sh base pad ...
# Times the prebuilt variants/<tag>.so libraries (tools/build_variant.sh) with quick_bench; tags listed in $TESTS
# also run the GPU parity tests first.
cd "$(dirname "$0")/.."
for tag in "$@"; do
  export OFDM_B200_LIB="$PWD/variants/$tag.so"
  echo "=== [$tag]"
  case " $TESTS " in *" $tag "*) python -m pytest tests -m gpu -x -q 2>&1 | tail -2;; esac
  for rep in 1 ${REPS:+2}; do
  python tools/quick_bench.py ${BENCH_SET:-all} 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    c = d.get('clocks') or {}
    print('%-12s rx %.3f ms %6.0f GB/s | tx %.3f ms %6.0f GB/s | ber %.2e locked %.3f | sm %s MHz %s' % (d['name'], d['rx_ms'], d['rx_GBs'], d['tx_ms'], d['tx_GBs'], d['ber'], d['locked'], c.get('sm_mhz'), ','.join(c.get('reasons', [])) or '-'))
"
  done
done
