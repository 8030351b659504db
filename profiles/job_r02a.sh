#!/bin/sh
# r02a (GPU box): parity of the new k_sched_tile / k_sor_ring + A/B of their knobs on configs[1]
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > $O/r02a_tests.log 2>&1
tail -5 $O/r02a_tests.log
b() { # tag, env...
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu > $O/r02a_$tag.json 2> $O/r02a_$tag.err
  python - "$O/r02a_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b new X=1
b old OB_SOR_RING=0 OB_SCHED_TILE=0
b ring_only OB_SCHED_TILE=0
b sched4 OB_SCHED_TILE=4
b sched16 OB_SCHED_TILE=16
b l2_40 OB_SOR_L2MB=40
b l2_60 OB_SOR_L2MB=60
b l2_100 OB_SOR_L2MB=100
b l2_1000 OB_SOR_L2MB=1000
b grid_740 OB_GRID_SOR=740
b grid_342 OB_GRID_SOR=342
b c4_new X=1 --config 4
b c4_old OB_SOR_RING=0 OB_SCHED_TILE=0 --config 4
b c3_new X=1 --config 3
b c3_ring_sched4 OB_SCHED_TILE=4 --config 3
b c3_old OB_SOR_RING=0 --config 3
