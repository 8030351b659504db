#!/bin/sh
# r02h (GPU box): k_prep split in two launches, schedule on a second stream beside the second half
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_batch or sampled or tile_width or lane_per_world" > $O/r02h_tests.log 2>&1
tail -4 $O/r02h_tests.log
OB_PREP_SPLIT=1 timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden" > $O/r02h_tests_split.log 2>&1
tail -3 $O/r02h_tests_split.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02h_$tag.json 2> $O/r02h_$tag.err
  python - "$O/r02h_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "sum %.3f"%sum(v["ms"] for v in k.values()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b split X=1
b nosplit OB_PREP_SPLIT=0
b split_prep8 OB_PREP_TILE=8
b c4_split X=1 --config 4
b c4_nosplit OB_PREP_SPLIT=0 --config 4
b c3_split X=1 --config 3
b c3_nosplit OB_PREP_SPLIT=0 --config 3
