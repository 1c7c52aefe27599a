F="--no-cpu-baseline --no-other-configs --no-multi-gpu-records --no-dense-boundary --steps 10 --warmup 3 --no-parity"
pick='import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d["value"], d["ms_per_step"], d["clocks"]["sm_mhz"])'
for h in 0 2 18 10 4 1 0 2; do echo HINTS $h; PIO_L2_HINTS=$h python bench.py $F 2>/dev/null | python -c "$pick"; done
