set -e
timeout 300 python -m pytest tests/test_gpu_host.py -m gpu -x -q 2>&1 | tail -5
D=/tmp/cfg2
python tools/make_config.py 2 $D > /dev/null
cd $D
for i in 1 2; do
  python - <<'PY'
import subprocess, time, os
t=time.time()
r=subprocess.run(["/root/repo/pedigreesim_b200/bin/pedigreesim","x.par"],env=dict(os.environ,PSIM_HOST_TIMING="1"),capture_output=True,text=True)
print("host program wall %.2f s"%(time.time()-t), r.stderr.strip())
PY
done
ls -la out* | awk '{print $5, $9}'
D2=/tmp/cfg2small
python /root/repo/tools/make_config.py 2 $D2 --popsize 250 > /dev/null
cd $D2
python - <<'PY'
import subprocess, time
t=time.time()
subprocess.run(["/root/repo/oracle/psim_oracle_cli","x.par","--philox"],capture_output=True)
print("oracle CLI (jar twin, popsize 250 of 10000) wall %.2f s"%(time.time()-t))
PY
