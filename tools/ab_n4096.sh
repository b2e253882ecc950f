#!/bin/bash
# A/B at N = 4096: per-lane stores (default) vs bulk rows through the async proxy, per modulation (run-time switches)
cd "$(dirname "$0")/.."
echo "== parity (default library)"; python -m pytest tests/test_gpu_parity.py tests/test_gpu_kernels.py -m gpu -x -q 2>&1 | tail -1
for rep in 1 2; do
echo "== default"; tools/qb.sh ab | grep -E "N4096|LTE"
echo "== bulk rows at N=4096 (OFDM_BULK_MIN_N=OFDM_BULK16_MIN_N=OFDM_BULK4_MIN_N=4096)"
OFDM_BULK_MIN_N=4096 OFDM_BULK16_MIN_N=4096 OFDM_BULK4_MIN_N=4096 tools/qb.sh ab | grep -E "N4096"
done
