#!/bin/sh
O=gpurun_out; mkdir -p $O
for gs in 444 592 740 1024; do
  OB_GRID_SOR=$gs python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_gs${gs}_c2.json 2> $O/bench_gs${gs}_c2.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_gs*_c2.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
