set -e
runhost() { # dir
python - "$1" <<'PY'
import subprocess, time, os, sys
d=sys.argv[1]
for i in range(2):
    t=time.time()
    r=subprocess.run(["/root/repo/pedigreesim_b200/bin/pedigreesim","x.par"],cwd=d,env=dict(os.environ,PSIM_HOST_TIMING="1"),capture_output=True,text=True)
    out=sum(os.path.getsize(os.path.join(d,f)) for f in os.listdir(d) if f.startswith("out"))
    print("host program run %d: wall %.2f s, %.2f GB of output files; %s"%(i, time.time()-t, out/1e9, r.stderr.strip()))
PY
}
runoracle() { # dir label
python - "$1" "$2" <<'PY'
import subprocess, time, sys
t=time.time()
subprocess.run(["/root/repo/oracle/psim_oracle_cli","x.par"],cwd=sys.argv[1],capture_output=True)
print("oracle CLI (C++ twin of the jar, 1 core) %s: wall %.2f s"%(sys.argv[2], time.time()-t))
PY
}
echo "== configs[1] full size (tetraploid F1 10000, 12 x 2000 loci, .gen)"
python tools/make_config.py 2 /tmp/c2 > /dev/null; runhost /tmp/c2
python tools/make_config.py 2 /tmp/c2s --popsize 250 > /dev/null; runoracle /tmp/c2s "popsize 250 of 10000"
echo "== configs[2] reduced (hexaploid F1 5000 of 100000, 21 x 1000 of 10000 loci, no .gen: founder alleles + all-allele doses)"
python tools/make_config.py 3 /tmp/c3 --popsize 5000 --loci 1000 > /dev/null; runhost /tmp/c3
python tools/make_config.py 3 /tmp/c3s --popsize 250 --loci 1000 > /dev/null; runoracle /tmp/c3s "popsize 250 of 5000"
echo "== configs[0] (diploid F2 of 200, 1 chromosome, 50 loci)"
python tools/make_config.py 1 /tmp/c1 > /dev/null; runhost /tmp/c1; runoracle /tmp/c1 "full size"
