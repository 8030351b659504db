#!/bin/sh
# r02s (GPU box): k_sor_lane (one lane per world) as the default for batches of many tiny worlds + its parity test on the golden scenes; whole GPU suite,
# bench of configs[1..3], launch list of configs[2]
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rs > $O/r02s_tests.log 2>&1
tail -4 $O/r02s_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02s_$tag.json 2> $O/r02s_$tag.err
  python - "$O/r02s_$tag.json" "$tag" <<'PY'
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
b c3_nolane OB_SOR_LANE=0 --config 3
b c4_lane OB_SOR_LANE=1 --config 4
ncu --metrics gpu__time_duration.sum --clock-control none -s 1200 -c 140 --csv --log-file $O/launches_r02s_c3.csv \
    python bench.py --config 3 --steps 12 --warmup 3 --no-cpu --no-other > $O/ncu_r02s_c3.log 2>&1
python - <<'PY'
import csv,collections
rows=[r for r in csv.reader(open("gpurun_out/launches_r02s_c3.csv")) if len(r)>5 and r[0].isdigit()]
agg=collections.defaultdict(list)
for r in rows:
    name=r[4].split("(")[0][:40]
    try: agg[name].append(float(r[-1]))
    except: pass
for k,v in agg.items(): print(k, len(v), "avg us %.1f"%(sum(v)/len(v)/ (1000.0 if max(v)>100000 else 1.0)))
PY
