#!/bin/sh
# r02m (GPU box): whole GPU suite after the shared broadphase function (filter / key split, grouped ranking); bench of configs[1..3] +
# SOR tile-width A/B with the current ring; ncu of k_collide (config 2) and of k_prep / k_collide_tile on configs[2] and [3]
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02m_tests.log 2>&1
tail -5 $O/r02m_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02m_$tag.json 2> $O/r02m_$tag.err
  python - "$O/r02m_$tag.json" "$tag" <<'PY'
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
b c2_tile16 OB_TILE=16
b c2_tile16_d3 OB_TILE=16 OB_RING_DEPTH=3
b c3 X=1 --config 3
b c4 X=1 --config 4
b c4_tile16 OB_TILE=16 --config 4
D=ode-0.12_b200/lib/driver_b200_single
cap() {   # tag kernel-regex skip scene worlds extra-args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02m_$1 \
      $D --scene $4 --worlds $5 $6 --steps 6 --settle 100 --mode batch --time > $O/ncu_r02m_$1.log 2>&1
  ncu -i $O/prof_r02m_$1.ncu-rep --page raw --csv > $O/raw_r02m_$1.csv 2>/dev/null
  ncu -i $O/prof_r02m_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02m_$1.csv.gz
  rm -f $O/prof_r02m_$1.ncu-rep
}
cap k_collide k_collide 102 stack32 4096 "--contacts-cap 192"
cap c4_k_prep k_prep 204 ragdoll 16384 "--contacts-cap 160"
cap c4_k_collide k_collide 102 ragdoll 16384 "--contacts-cap 160"
cap c3_k_prep k_prep 204 buggy_terrain256 65536 "--contacts-cap 48"
cap c3_k_collide_tile k_collide_tile 102 buggy_terrain256 65536 "--contacts-cap 48"
tail -3 $O/ncu_r02m_c3_k_prep.log
