#!/bin/sh
# r02o (GPU box): decoupled narrowphase (k_broad -> k_narrow -> k_contacts): whole GPU suite, A/B against the fused kernels on
# configs[1..3], ncu of the three new kernels on configs[1] and [2]
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02o_tests.log 2>&1
tail -5 $O/r02o_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02o_$tag.json 2> $O/r02o_$tag.err
  python - "$O/r02o_$tag.json" "$tag" <<'PY'
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
b c2_fused OB_COLLIDE_FUSED=1
b c3 X=1 --config 3
b c3_fused OB_COLLIDE_FUSED=1 --config 3
b c4 X=1 --config 4
b c4_fused OB_COLLIDE_FUSED=1 --config 4
D=ode-0.12_b200/lib/driver_b200_single
cap() {   # tag kernel-regex skip scene worlds extra-args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02o_$1 \
      $D --scene $4 --worlds $5 $6 --steps 6 --settle 100 --mode batch --time > $O/ncu_r02o_$1.log 2>&1
  ncu -i $O/prof_r02o_$1.ncu-rep --page raw --csv > $O/raw_r02o_$1.csv 2>/dev/null
  ncu -i $O/prof_r02o_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02o_$1.csv.gz
  rm -f $O/prof_r02o_$1.ncu-rep
}
cap k_broad "k_broad" 102 stack32 4096 "--contacts-cap 192"
cap k_narrow k_narrow 102 stack32 4096 "--contacts-cap 192"
cap k_contacts k_contacts 102 stack32 4096 "--contacts-cap 192"
cap c3_k_narrow k_narrow 102 buggy_terrain256 65536 "--contacts-cap 48"
cap c3_k_broad_tile k_broad_tile 102 buggy_terrain256 65536 "--contacts-cap 48"
