#!/bin/sh
# r02j (GPU box): the driver's two bench commands with wall-clock times; launch list + ncu captures of the five kernels of configs[1]
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q > $O/r02j_tests.log 2>&1
tail -8 $O/r02j_tests.log
nproc
t0=$(date +%s)
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02j_bench_ref.json 2> $O/r02j_bench_ref.err
t1=$(date +%s); echo "reference arm wall $((t1-t0)) s"
tail -c 700 $O/r02j_bench_ref.json; echo
python bench.py --steps 20 --warmup 5 > $O/r02j_bench.json 2> $O/r02j_bench.err
t2=$(date +%s); echo "b200 arm wall $((t2-t1)) s"
tail -3 $O/r02j_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02j_bench.json").read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"value %.3g"%d["value"],"e2e %.3g"%d["e2e"]["value"],"cpu",d.get("cpu_baseline",{}).get("value"), "kernels", {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()})
for k,v in d.get("other_configs",{}).items():
    print(k, v and {x:v.get(x) for x in ("ms_per_step","value","e2e","whole_step_frac_of_hbm_peak")})
PY
for ch in 2 4; do
  OB_CHUNKS=$ch python bench.py --steps 30 --warmup 3 --no-cpu --no-other > $O/r02j_chunks$ch.json 2> $O/r02j_chunks$ch.err
  python -c "import json;d=json.loads(open('$O/r02j_chunks$ch.json').read().strip().splitlines()[-1]);print('chunks $ch ms/step',d['ms_per_step'])"
done
# launch list of the default bench command (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1800 -c 300 --csv --log-file $O/launches_r02j_c2.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu --no-other > $O/ncu_r02j_c2.log 2>&1
D=ode-0.12_b200/lib/driver_b200_single
for k in k_sor_ring k_collide k_prep k_sched_tile k_post; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02j_$k \
      $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02j_$k.log 2>&1
  ncu -i $O/prof_r02j_$k.ncu-rep --page raw --csv > $O/raw_r02j_$k.csv 2>/dev/null
  ncu -i $O/prof_r02j_$k.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_r02j_$k.csv.gz
  rm -f $O/prof_r02j_$k.ncu-rep
done
ls -la $O | tail -15
