#!/bin/sh
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "pistons or pus" > $O/joint_tests_s4e.log 2>&1; tail -4 $O/joint_tests_s4e.log
python bench.py --steps 30 --warmup 3 > $O/bench_s4e_c2.json 2> $O/bench_s4e_c2.err
python bench.py --config 1 --steps 200 --warmup 5 > $O/bench_s4e_c1.json 2> $O/bench_s4e_c1.err
python bench.py --config 1 --impl reference --steps 200 --warmup 5 > $O/bench_s4e_c1_ref.json 2> $O/bench_s4e_c1_ref.err
tail -n 3 $O/bench_s4e_c1.err $O/bench_s4e_c2.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_s4e_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, d["n_gpus"], "%.3e" % d["value"], "%.4f ms" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], {k: round(v["ms"], 4) for k, v in r.get("kernels", {}).items()}, d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(f, "ERR", e)
PY
