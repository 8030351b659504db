#!/bin/sh
# r02g (GPU box): k_sor_reg (pipelined pass, register row buffers)
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or full_batch or sampled or tile_width" > $O/r02g_tests.log 2>&1
tail -4 $O/r02g_tests.log
b() {
  tag=$1; shift
  ENVS=""; ARGS=""
  for a in "$@"; do case "$a" in --*|[0-9]*) ARGS="$ARGS $a";; *) ENVS="$ENVS $a";; esac; done
  env $ENVS python bench.py $ARGS --steps 30 --warmup 3 --no-cpu --no-other > $O/r02g_$tag.json 2> $O/r02g_$tag.err
  python - "$O/r02g_$tag.json" "$tag" <<'PY'
import json,sys
try:
    d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    k=d["roofline"]["kernels"]
    print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
except Exception as e:
    print(sys.argv[2], "FAILED", e)
PY
}
b reg X=1
b ring OB_SOR_REG=0
b old OB_SOR_RING=0
b reg_t16 OB_TILE=16
b c4_reg X=1 --config 4
b c4_ring OB_SOR_REG=0 --config 4
b c3_reg OB_SOR_RING=1 --config 3
b c3_old X=1 --config 3
D=ode-0.12_b200/lib/driver_b200_single
for k in k_sor_reg; do
ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02g_$k \
    $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02g_$k.log 2>&1
ncu -i $O/prof_r02g_$k.ncu-rep --page raw --csv > $O/raw_r02g_$k.csv 2>/dev/null
ncu -i $O/prof_r02g_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02g_$k.csv.gz
rm -f $O/prof_r02g_$k.ncu-rep
done
