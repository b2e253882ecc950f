#!/bin/bash
# Run on the GPU box (gpurun): bench lines, launch list and full ncu captures for profiles/.
# Every ncu command is preceded by the same command without ncu exiting 0.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/ev
mkdir -p $O
python bench.py --steps 50 --warmup 5 > $O/bench_n64.json 2> $O/bench_n64.err || { echo "bench failed"; tail -5 $O/bench_n64.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > $O/bench_reference.json 2> $O/bench_reference.err || echo "reference arm failed"
python bench.py --steps 3 --warmup 3 --no-cpu > $O/bench_short.json 2> $O/bench_short.err && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench_n64.csv \
      python bench.py --steps 3 --warmup 3 --no-cpu > $O/ncu_launches.log 2>&1
python tools/prof_cases.py n64 lte n4096 > $O/prof_plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"rx_frames_kernel|tx_kernel|tx_wide_kernel" -s 2 -c 18 -o $O/prof_final -f \
      python tools/prof_cases.py n64 lte n4096 > $O/prof_ncu.log 2>&1
# acquisition kernels: N-tap kernel (N=64), overlap-save (LTE), general form (two synch symbols; hypothesis bank)
if [ -z "${SKIP_SYNC_NCU:-}" ]; then
python tools/bench_sync.py n64_utsa_stride1 lte_stride1 n64_m2_stride1 lte_bank3_stride1 --no-cpu > $O/prof_sync_plain.log 2>&1 && \
  ncu --set full --clock-control none -k regex:"rx_sync_search_(td8|os|tdg)" -c 40 -o $O/prof_sync -f \
      python tools/bench_sync.py n64_utsa_stride1 lte_stride1 n64_m2_stride1 lte_bank3_stride1 --no-cpu > $O/prof_sync_ncu.log 2>&1
fi
# gpurun brings back at most 64 MiB: keep the raw pages (what tools/publish_evidence.py reads) and drop the reports
for rep in prof_final prof_sync; do
  if [ -f $O/$rep.ncu-rep ]; then
    ncu -i $O/$rep.ncu-rep --page raw --csv > $O/${rep}_raw.csv 2> /dev/null
    rm -f $O/$rep.ncu-rep
  fi
done
python tools/sweep_nfft.py > $O/sweep_nfft.jsonl 2> $O/sweep_nfft.err
python bench.py --workload rx_lte20_64qam --steps 20 --warmup 3 --no-cpu > $O/bench_lte.json 2> $O/bench_lte.err
python bench.py --workload rx_lte20_streams --steps 10 --warmup 3 --no-cpu > $O/bench_lte_streams.json 2> $O/bench_lte_streams.err
python tools/bench_sync.py > $O/bench_sync_search.jsonl 2> $O/bench_sync_search.err
python tools/bench_blocks.py > $O/bench_blocks.jsonl 2> $O/bench_blocks.err
ls -la $O
