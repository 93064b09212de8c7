set -e
python tools/make_config.py 4 /tmp/c4 --popsize 50000 --loci 200 > /dev/null
cd /tmp/c4
python - <<'PY'
import subprocess, time, os
t=time.time()
r=subprocess.run(["/root/repo/pedigreesim_b200/bin/pedigreesim","x.par"],env=dict(os.environ,PSIM_HOST_TIMING="1"),capture_output=True,text=True)
out={f:os.path.getsize(f) for f in os.listdir(".") if f.startswith("out")}
print("host program: wall %.2f s; %s"%(time.time()-t, r.stderr.strip()))
print(r.stdout[-300:])
print(out)
# spot check: founder rows of the last F8 line are homozygous at most loci and P1/P2 columns are 0,1 / 2,3
with open("out_founderalleles.dat") as f:
    head=f.readline().rstrip("\n").split("\t")
    row=f.readline().rstrip("\n").split("\t")
print(len(head), head[:6], row[:6], head[-2:], row[-2:])
PY
