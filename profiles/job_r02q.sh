#!/bin/sh
# r02q (GPU box): lock-step BVH walk / triangle test in the sphere- and box-trimesh colliders; scheduler reverted to r02o's;
# mesh + golden parity subset, bench of configs[1..3], launch list of configs[2]
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -x -k "terrain or buggy or golden or raycast or KAT or kat or trimesh or tile or sampled" > $O/r02q_tests.log 2>&1
tail -4 $O/r02q_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02q_$tag.json 2> $O/r02q_$tag.err
  python - "$O/r02q_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "sum %.3f"%sum(v["ms"] for v in k.values()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b c2 X=1
b c3 X=1 --config 3
b c4 X=1 --config 4
ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 120 --csv --log-file $O/launches_r02q_c3.csv \
    python bench.py --config 3 --steps 12 --warmup 3 --no-cpu --no-other > $O/ncu_r02q_c3.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_r02q_c3.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    name=r[4].split("(")[0][:40]
    try: agg[name].append(float(r[-1]))
    except: pass
for k,v in agg.items(): print(k, len(v), "avg us %.1f"%(sum(v)/len(v)/ (1000.0 if max(v)>100000 else 1.0)))
PY
