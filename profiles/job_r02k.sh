#!/bin/sh
# r02k (GPU box): nested-space scenes after the shared-memory attribute fix; box-box with the straight-line axis scan + triangular
# candidate enumeration: parity subset, bench of configs[1..3], ncu of k_collide
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "nested" > $O/r02k_tests_nested.log 2>&1
tail -3 $O/r02k_tests_nested.log
timeout 1500 python -m pytest tests -m gpu -q -x -k "not nested" > $O/r02k_tests.log 2>&1
tail -3 $O/r02k_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02k_$tag.json 2> $O/r02k_$tag.err
  python - "$O/r02k_$tag.json" "$tag" <<'PY'
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
D=ode-0.12_b200/lib/driver_b200_single
for k in k_collide; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02k_$k \
      $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02k_$k.log 2>&1
  ncu -i $O/prof_r02k_$k.ncu-rep --page raw --csv > $O/raw_r02k_$k.csv 2>/dev/null
  ncu -i $O/prof_r02k_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02k_$k.csv.gz
  rm -f $O/prof_r02k_$k.ncu-rep
done
