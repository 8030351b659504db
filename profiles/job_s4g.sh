#!/bin/sh
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests -m gpu -q -k "kinematic or dif or raycyl" > $O/tests_s4g.log 2>&1; tail -4 $O/tests_s4g.log
timeout 300 python -m pytest tests/test_api_probe.py tests/test_abi.py -q > $O/tests_s4g_cpu.log 2>&1; tail -2 $O/tests_s4g_cpu.log
python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_s4g_c2.json 2> $O/bench_s4g_c2.err
python - <<PY
import json
d = json.loads(open("$O/bench_s4g_c2.json").read().strip().splitlines()[-1])
print("%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()}, "whole", round(d["roofline"]["whole_step"]["achieved"]))
PY
