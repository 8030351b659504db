#!/bin/sh
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "cyl" > $O/cyl_tests_s4f.log 2>&1; tail -4 $O/cyl_tests_s4f.log
python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_s4f_c2.json 2> $O/bench_s4f_c2.err
python - <<PY
import json
d = json.loads(open("$O/bench_s4f_c2.json").read().strip().splitlines()[-1])
print("%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
PY
