#!/bin/sh
# r02i (GPU box): the whole GPU suite, then the driver's two bench commands with wall-clock times
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -x -q > $O/r02i_tests.log 2>&1
tail -5 $O/r02i_tests.log
nproc
( /usr/bin/time -f "real %e s" python bench.py --impl reference --steps 20 --warmup 5 > $O/r02i_bench_ref.json 2> $O/r02i_bench_ref.err ) 2>&1 | grep real
tail -c 600 $O/r02i_bench_ref.json
( /usr/bin/time -f "real %e s" python bench.py --steps 20 --warmup 5 > $O/r02i_bench.json 2> $O/r02i_bench.err ) 2>&1 | grep real
tail -3 $O/r02i_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02i_bench.json").read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"value %.3g"%d["value"],"e2e %.3g"%d["e2e"]["value"],"cpu",d.get("cpu_baseline",{}).get("value"))
for k,v in d.get("other_configs",{}).items():
    print(k, v and {x:v[x] for x in ("ms_per_step","value","e2e","whole_step_frac_of_hbm_peak")})
PY
