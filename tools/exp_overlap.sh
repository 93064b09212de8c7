#!/bin/bash
# The pipelined bench step: k_emit_rows at 48 registers with 512 / 256 compute threads x k_meiosis at 64 / 56 registers
mkdir -p gpurun_out
{
for lib in build/libpsim_F2.so build/libpsim_C.so build/libpsim_F.so build/libpsim_tune.so build/libpsim_F2.so; do
    echo "== $lib"
    PSIM_TUNING_LIB=$lib timeout 300 python bench.py --steps 40 --no-cpu 2>&1 | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.0f ms_per_step %.3f packed %.0f (%.3f ms) kernel_ms %s roofline %.3f sweep %s' % (d['value'], d['ms_per_step'], d['packed_tables']['value'], d['packed_tables']['ms_per_step'], json.dumps({k: round(v, 3) for k, v in d['kernel_ms'].items()}), d['roofline']['frac'], json.dumps({k: round(d['replicate_sweep'][k],3) for k in ('simulate_ms','emit_kernel_ms','emit_all_kernels_ms')})))"
done
} > gpurun_out/r2_overlap5.log 2>&1
cat gpurun_out/r2_overlap5.log
