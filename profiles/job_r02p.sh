#!/bin/sh
# r02p (GPU box): k_sched_tile as a two-chain software pipeline (shuffle of epoch e + 1 beside the levels of epoch e), k_contacts<G>,
# separate tile width for the first half of k_prep: whole GPU suite, A/B on configs[1..3]
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02p_tests.log 2>&1
tail -5 $O/r02p_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02p_$tag.json 2> $O/r02p_$tag.err
  python - "$O/r02p_$tag.json" "$tag" <<'PY'
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
b c2_p1t8 OB_PREP_TILE1=8
b c2_p1t16 OB_PREP_TILE1=16
b c2_sched8 OB_SCHED_TILE=8
b c3 X=1 --config 3
b c3_p1t8 OB_PREP_TILE1=8 --config 3
b c4 X=1 --config 4
b c4_p1t4 OB_PREP_TILE1=4 --config 4
b c4_p1t16 OB_PREP_TILE1=16 --config 4
b c4_sched8 OB_SCHED_TILE=8 --config 4
D=ode-0.12_b200/lib/driver_b200_single
cap() {   # tag kernel-regex skip scene worlds extra-args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02p_$1 \
      $D --scene $4 --worlds $5 $6 --steps 6 --settle 100 --mode batch --time > $O/ncu_r02p_$1.log 2>&1
  ncu -i $O/prof_r02p_$1.ncu-rep --page raw --csv > $O/raw_r02p_$1.csv 2>/dev/null
  ncu -i $O/prof_r02p_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02p_$1.csv.gz
  rm -f $O/prof_r02p_$1.ncu-rep
}
cap k_sched_tile k_sched_tile 102 stack32 4096 "--contacts-cap 192"
cap c4_k_prep2 k_prep 205 ragdoll 16384 "--contacts-cap 160"
