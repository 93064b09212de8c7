#!/bin/bash
# The pipelined bench step with (a) fewer compute threads in k_emit_rows, so that more meiosis blocks fit next to it,
# (b) the meiosis kernels on a lowest-priority stream (PSIM_SIM_PRIO=-1: on the context's stream as before)
mkdir -p gpurun_out
{
for lib in build/libpsim_tune.so build/libpsim_c384g1.so build/libpsim_c256g1.so build/libpsim_c256g2.so; do
  for prio in -1 0; do
    echo "== $lib PSIM_SIM_PRIO=$prio"
    PSIM_SIM_PRIO=$prio PSIM_TUNING_LIB=$lib timeout 300 python bench.py --steps 20 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f kernel_ms %s roofline %.3f e2e %.1f sweep %s' % (d['value'], d['ms_per_step'], json.dumps(d['kernel_ms']), d['roofline']['frac'], d['e2e']['value'], json.dumps({k: d['replicate_sweep'].get(k) for k in ('simulate_ms','emit_kernel_ms','emit_all_kernels_ms')})))"
  done
done
echo "== sweep, shipped library"
timeout 300 python tools/prof_sweep.py
} > gpurun_out/r2_overlap.log 2>&1
timeout 900 python -m pytest tests/test_gpu_emit.py tests/test_gpu_configs.py tests/test_gpu_meiosis.py -x -q -m gpu > gpurun_out/r2_overlap_tests.log 2>&1
tail -3 gpurun_out/r2_overlap_tests.log
cat gpurun_out/r2_overlap.log
