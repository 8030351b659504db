#!/bin/sh
# r02e (GPU box): k_sor_pair (two lanes per row), new parity scenes / KATs / libm on the device, build without --split-compile
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_reference_kats.py tests/test_abi.py tests/test_gpu_parity.py -m gpu -x -q -k "kats or libm or golden or bodyflags or autodisable or contactmodes or varmaxc or hinges or buggy or ragdoll or full_batch or sampled or tile_width" > $O/r02e_tests.log 2>&1
tail -6 $O/r02e_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02e_$tag.json 2> $O/r02e_$tag.err
  python - "$O/r02e_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b pair X=1
b pair_d4 OB_PAIR_DEPTH=4
b ring OB_SOR_PAIR=0
b old OB_SOR_RING=0
b pair_sched16 OB_SCHED_TILE=16
b c4_pair X=1 --config 4
b c4_ring OB_SOR_PAIR=0 --config 4
b c3_pair X=1 --config 3
b c3_ring OB_SOR_PAIR=0 --config 3
b c3_old OB_SOR_RING=0 --config 3
D=ode-0.12_b200/lib/driver_b200_single
for k in k_sor_pair; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02e_$k \
    $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02e_$k.log 2>&1
ncu -i $O/prof_r02e_$k.ncu-rep --page raw --csv > $O/raw_r02e_$k.csv 2>/dev/null
ncu -i $O/prof_r02e_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02e_$k.csv.gz
rm -f $O/prof_r02e_$k.ncu-rep
done
