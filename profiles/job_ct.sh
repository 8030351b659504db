#!/bin/sh
# experiment: CTA width of k_collide on config 2 / 4
O=gpurun_out; mkdir -p $O
for ct in 64 96 128; do
  OB_COLLIDE_THREADS=$ct python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_ct${ct}_c2.json 2> $O/bench_ct${ct}_c2.err
  OB_COLLIDE_THREADS=$ct python bench.py --config 4 --steps 30 --warmup 3 --no-cpu > $O/bench_ct${ct}_c4.json 2> $O/bench_ct${ct}_c4.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_ct*_c*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f, "%.4f ms" % d["ms_per_step"], {k: round(v["ms"], 4) for k, v in d["roofline"]["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
