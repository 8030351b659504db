#!/bin/sh
# r02t (GPU box): k_sor_lane wave size (row records of the worlds in flight vs the 126 MB L2) on configs[2]
O=gpurun_out
mkdir -p $O
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02t_$tag.json 2> $O/r02t_$tag.err
  python - "$O/r02t_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "sum %.3f"%sum(v["ms"] for v in k.values()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b c3_g1024 OB_SOR_LANE_GRID=1024 --config 3
b c3_g683 OB_SOR_LANE_GRID=683 --config 3
b c3_g512 OB_SOR_LANE_GRID=512 --config 3
b c3_g342 OB_SOR_LANE_GRID=342 --config 3
b c3_g256 OB_SOR_LANE_GRID=256 --config 3
