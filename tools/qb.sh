#!/bin/bash
# pretty-printed quick_bench:  tools/qb.sh <set>
cd "$(dirname "$0")/.."
python tools/quick_bench.py "$@" 2>&1 | python -c "
import sys, json
for l in sys.stdin:
    try: d = json.loads(l)
    except Exception: print(l.strip()); continue
    c = d.get('clocks') or {}
    print('%-12s rx %.3f ms %6.0f GB/s | tx %.3f ms %6.0f GB/s | ber %.2e locked %.3f | sm %s MHz %s' % (d['name'], d['rx_ms'], d['rx_GBs'], d['tx_ms'], d['tx_GBs'], d['ber'], d['locked'], c.get('sm_mhz'), ','.join(c.get('reasons', [])) or '-'))
"
