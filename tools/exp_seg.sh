#!/bin/bash
mkdir -p gpurun_out
{
for seg in 0 28672 24576 20480 16384 12288 8192; do
  echo "== PSIM_SEG_MAX=$seg"
  PSIM_SEG_MAX=$seg PSIM_ROWS_TRACE=1 PSIM_TUNING_LIB=build/libpsim_tune.so timeout 300 python tools/prof_sweep.py 2>&1 | tail -2 | cut -c1-330
done
} > gpurun_out/r2_seg.log 2>&1
cat gpurun_out/r2_seg.log
