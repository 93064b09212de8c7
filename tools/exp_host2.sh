#!/bin/bash
# The host program `pedigreesim x.par` on BASELINE configs[1] (full size) and configs[0], with the time of its phases
mkdir -p gpurun_out
{
for cfg in 2 1; do
  rm -rf /tmp/hc$cfg; python tools/make_config.py $cfg /tmp/hc$cfg > /dev/null
  for run in 0 1; do
    echo "== make_config $cfg, run $run"
    ( cd /tmp/hc$cfg && rm -f out* && python3 - <<PY
import os, subprocess, time
t = time.time()
r = subprocess.run(["$GRAFT_REPO_ROOT/pedigreesim_b200/bin/pedigreesim", "x.par"], env=dict(os.environ, PSIM_HOST_TIMING="1"), capture_output=True, text=True)
print("wall %.2f s; %s" % (time.time() - t, r.stderr.strip()))
print("%.2f GB of output files" % (sum(os.path.getsize(f) for f in os.listdir(".") if f.startswith("out")) / 1e9))
PY
    )
  done
done
} > gpurun_out/r2_host_program.txt 2>&1
cat gpurun_out/r2_host_program.txt
