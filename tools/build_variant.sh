#!/bin/bash
# usage: tools/build_variant.sh <tag> "<nvcc -D flags>"
# Cross-compiles a tuning variant of the library HERE (no GPU needed) into variants/<tag>.so so that a GPU call only
# has to time it:  OFDM_B200_LIB=variants/<tag>.so python tools/quick_bench.py ...
set -e
ROOT="$(cd "$(dirname "$0")/.." && pwd)"
tag="$1"; defs="$2"
mkdir -p "$ROOT/variants"
cd "$ROOT/lte-gnu-radio-code_b200/csrc"
make -j8 OBJDIR="/tmp/ofdm_variant_$tag" OUT="$ROOT/variants/$tag.so" EXTRA="$defs" > "/tmp/ofdm_variant_$tag.log" 2>&1 || { tail -20 "/tmp/ofdm_variant_$tag.log"; exit 1; }
echo "$tag [$defs] spills: $(grep -h 'spill' /tmp/ofdm_variant_$tag/*.ptxas.log | grep -vc ' 0 bytes spill stores')"
