#!/bin/sh
# r02c (GPU box): pipelined branch-free k_sor_ring<G, D>: parity subset, A/B on configs[1], [2], [3], ncu of the kernel
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or tile_width or full_batch or sampled_worlds" > $O/r02c_tests.log 2>&1
tail -4 $O/r02c_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02c_$tag.json 2> $O/r02c_$tag.err
  python - "$O/r02c_$tag.json" "$tag" <<'PY'
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
b old OB_SOR_RING=0
b d4 OB_RING_DEPTH=4
b d6 OB_RING_DEPTH=6
b tile16 OB_TILE=16
b tile16_d4 OB_TILE=16 OB_RING_DEPTH=4
b tile4 OB_TILE=4
b c4_new X=1 --config 4
b c4_old OB_SOR_RING=0 --config 4
b c4_t16 OB_TILE=16 --config 4
b c3_new X=1 --config 3
b c3_old OB_SOR_RING=0 --config 3
b c3_t8 OB_TILE=8 --config 3
D=ode-0.12_b200/lib/driver_b200_single
ncu --set full --clock-control none --import-source on -k regex:k_sor_ring -s 305 -c 1 -f -o $O/prof_r02c_k_sor_ring \
    $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02c_k_sor_ring.log 2>&1
ncu -i $O/prof_r02c_k_sor_ring.ncu-rep --page raw --csv > $O/raw_r02c_k_sor_ring.csv 2>/dev/null
ncu -i $O/prof_r02c_k_sor_ring.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02c_k_sor_ring.csv.gz
rm -f $O/prof_r02c_k_sor_ring.ncu-rep
