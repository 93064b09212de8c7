#!/bin/bash
# The host program `pedigreesim x.par` on BASELINE configs[1] (full size) and configs[0], with the time of its phases
mkdir -p gpurun_out
{
for cfg in 2 1; do
  rm -rf /tmp/hc$cfg; python tools/make_config.py $cfg /tmp/hc$cfg > /dev/null
  for run in 0 1; do
    echo "== make_config $cfg, run $run"
    ( cd /tmp/hc$cfg && rm -f out* && t0=$(date +%s.%N) && PSIM_HOST_TIMING=1 $GRAFT_REPO_ROOT/pedigreesim_b200/bin/pedigreesim x.par 2>&1 | tail -2; echo "wall $(echo "$(date +%s.%N) - $t0" | bc) s"; du -cb out* | tail -1 )
  done
done
} > gpurun_out/r2_host_program.txt 2>&1
cat gpurun_out/r2_host_program.txt
