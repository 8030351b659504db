#!/bin/sh
# quick check on the GPU box: golden-trace parity + default bench (+ configs 3, 4 with QUICK_ALL=1)
O=gpurun_out; mkdir -p $O; TAG=${1:-q}
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > $O/quick_tests_$TAG.log 2>&1; tail -3 $O/quick_tests_$TAG.log
python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_${TAG}_c2.json 2> $O/bench_${TAG}_c2.err
if [ -n "$QUICK_ALL" ]; then
  python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O/bench_${TAG}_c3.json 2> $O/bench_${TAG}_c3.err
  python bench.py --config 4 --steps 30 --warmup 3 --no-cpu > $O/bench_${TAG}_c4.json 2> $O/bench_${TAG}_c4.err
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_${TAG}_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d.get("roofline", {})
        print(f, "%.3e" % d["value"], "%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in r.get("kernels", {}).items()}, "whole", round(r["whole_step"]["achieved"]))
    except Exception as e:
        print(f, "ERR", e)
PY
