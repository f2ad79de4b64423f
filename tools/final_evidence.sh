#!/bin/bash
# Refreshes the evidence under profiles/ (gpurun_out/ must stay below 64 MiB to travel back, hence two calls):
#   tools/final_evidence.sh bench   full 1-GPU bench line, reference arm, ncu launch list of the device-only bench
#   tools/final_evidence.sh full    ncu --set full of one launch group of every main kernel (after the plain run)
mkdir -p gpurun_out
BENCH="python bench.py --steps 2 --warmup 1 --device-only --no-cpu-baseline"
if [ "$1" = bench ]; then
  python bench.py --steps 20 --warmup 3 > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
  python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"
  $BENCH > gpurun_out/plain.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv $BENCH > gpurun_out/ncu1.log 2>&1
  echo "launches rc=$?"
else
  $BENCH > gpurun_out/plain.log 2>&1 && \
  ncu --set full --clock-control none --import-source on -k regex:"k_nr_|k_frame_|k_quantize" --launch-skip 16 --launch-count 12 \
      -o gpurun_out/prof_all -f $BENCH > gpurun_out/ncu2.log 2>&1
  echo "full rc=$?"; ls -la gpurun_out/prof_all.ncu-rep
fi
