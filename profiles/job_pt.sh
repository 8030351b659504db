#!/bin/sh
# experiment: k_prep tile width
O=gpurun_out; mkdir -p $O
for pt in 8 16 32; do
  OB_PREP_TILE=$pt python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_pt${pt}_c2.json 2> $O/bench_pt${pt}_c2.err
  OB_PREP_TILE=$pt python bench.py --config 4 --steps 30 --warmup 3 --no-cpu > $O/bench_pt${pt}_c4.json 2> $O/bench_pt${pt}_c4.err
  OB_PREP_TILE=$pt python bench.py --config 3 --steps 30 --warmup 3 --no-cpu > $O/bench_pt${pt}_c3.json 2> $O/bench_pt${pt}_c3.err
done
OB_PREP_TILE=32 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > $O/pt32_tests.log 2>&1; tail -2 $O/pt32_tests.log
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_pt*_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
