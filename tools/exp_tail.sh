# sweep of the two-round band plan of the row-image emission (tuning build: -DPSIM_TUNING)
export PSIM_TUNING_LIB=build/libpsim_tune.so
run() { echo "$@"; env "$@" python tools/prof_step.py --steps 5 2>&1 | tail -2; }
run PSIM_ROWS_UNITS=148
for pct in 15 25 35 50; do for w in 2 3 5; do run PSIM_ROWS_TAIL_PCT=$pct PSIM_ROWS_TAIL_WAVES=$w; done; done
