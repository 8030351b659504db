#!/bin/sh
# r02z (GPU box): the round's closing run -- whole GPU suite, the driver's two bench commands (reference arm first), launch list of the
# default bench command, ncu --set full of every kernel of configs[1] (and of the sweeps of configs[2], [3])
O=gpurun_out
mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -rs > $O/r02z_tests.log 2>&1
tail -8 $O/r02z_tests.log
nproc
t0=$(date +%s)
python bench.py --impl reference --steps 20 --warmup 5 > $O/r02z_bench_ref.json 2> $O/r02z_bench_ref.err
t1=$(date +%s); echo "reference arm wall $((t1-t0)) s"
python bench.py --steps 20 --warmup 5 > $O/r02z_bench.json 2> $O/r02z_bench.err
t2=$(date +%s); echo "b200 arm wall $((t2-t1)) s"
tail -3 $O/r02z_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r02z_bench.json").read().strip().splitlines()[-1])
r=json.loads(open("gpurun_out/r02z_bench_ref.json").read().strip().splitlines()[-1])
print("ms/step",d["ms_per_step"],"value %.3g"%d["value"],"e2e %.3g"%d["e2e"]["value"],"ref %.3g"%r["value"],"e2e ratio %.1f"%(d["e2e"]["value"]/r["value"]), "clocks", d["clocks"], "kernels", {k:round(v["ms"],3) for k,v in d["roofline"]["kernels"].items()})
for k,v in d.get("other_configs",{}).items():
    print(k, v and {x:v.get(x) for x in ("ms_per_step","value","e2e","whole_step_frac_of_hbm_peak")})
PY
ncu --metrics gpu__time_duration.sum --clock-control none -s 2400 -c 320 --csv --log-file $O/launches_r02z_c2.csv \
    python bench.py --steps 30 --warmup 3 --no-cpu --no-other > $O/ncu_r02z_c2.log 2>&1
D=ode-0.12_b200/lib/driver_b200_single
cap() {   # tag kernel-regex skip scene worlds extra-args
  ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02z_$1 \
      $D --scene $4 --worlds $5 $6 --steps 6 --settle 150 --mode batch --time > $O/ncu_r02z_$1.log 2>&1
  ncu -i $O/prof_r02z_$1.ncu-rep --page raw --csv > $O/raw_r02z_$1.csv 2>/dev/null
  rm -f $O/prof_r02z_$1.ncu-rep
}
cap k_sor_ring k_sor_ring 152 stack32 4096 "--contacts-cap 192"
cap k_broad "k_broad" 152 stack32 4096 "--contacts-cap 192"
cap k_narrow k_narrow 152 stack32 4096 "--contacts-cap 192"
cap k_contacts k_contacts 152 stack32 4096 "--contacts-cap 192"
cap k_prep1 k_prep 304 stack32 4096 "--contacts-cap 192"
cap k_prep2 k_prep 305 stack32 4096 "--contacts-cap 192"
cap k_sched_tile k_sched_tile 152 stack32 4096 "--contacts-cap 192"
cap k_post k_post 152 stack32 4096 "--contacts-cap 192"
cap c3_k_sor_lane k_sor_lane 152 buggy_terrain256 65536 "--contacts-cap 48"
cap c4_k_sor_ring k_sor_ring 152 ragdoll 16384 "--contacts-cap 160"
ls $O | grep r02z | wc -l
