#!/bin/sh
# r02d (GPU box): ring v3 (ping-pong, opaque lane), k_sched_tile with flat serial loops, k_collide round-robin narrowphase
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or tile_width or full_batch or sampled_worlds or lane_per_world" > $O/r02d_tests.log 2>&1
tail -4 $O/r02d_tests.log
OB_SCHED_TILE=8 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden and single or full_batch or sampled_worlds" > $O/r02d_tests_sched8.log 2>&1
tail -3 $O/r02d_tests_sched8.log
OB_SCHED_TILE=4 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden and single" > $O/r02d_tests_sched4.log 2>&1
tail -3 $O/r02d_tests_sched4.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02d_$tag.json 2> $O/r02d_$tag.err
  python - "$O/r02d_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b ring X=1
b old OB_SOR_RING=0
b sched8 OB_SCHED_TILE=8 OB_SOR_RING=0
b sched16 OB_SCHED_TILE=16 OB_SOR_RING=0
b sched4 OB_SCHED_TILE=4 OB_SOR_RING=0
b ct64 OB_COLLIDE_THREADS=64 OB_SOR_RING=0
b ct96 OB_COLLIDE_THREADS=96 OB_SOR_RING=0
b c4_ring X=1 --config 4
b c4_sched8 OB_SCHED_TILE=8 OB_SOR_RING=0 --config 4
b c4_sched4 OB_SCHED_TILE=4 OB_SOR_RING=0 --config 4
b c3_sched4 OB_SCHED_TILE=4 OB_SOR_RING=0 --config 3
b c3_sched2 OB_SCHED_TILE=2 OB_SOR_RING=0 --config 3
b c3_old OB_SOR_RING=0 --config 3
D=ode-0.12_b200/lib/driver_b200_single
for k in k_sched_tile k_collide; do
OB_SCHED_TILE=8 ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02d_$k \
    $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02d_$k.log 2>&1
ncu -i $O/prof_r02d_$k.ncu-rep --page raw --csv > $O/raw_r02d_$k.csv 2>/dev/null
ncu -i $O/prof_r02d_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02d_$k.csv.gz
rm -f $O/prof_r02d_$k.ncu-rep
done
