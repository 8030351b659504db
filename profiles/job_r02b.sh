#!/bin/sh
# r02b (GPU box): ncu --set full of the new step kernels on configs[1] (4096 x stack32), raw metrics + source page
O=gpurun_out
mkdir -p $O
D=ode-0.12_b200/lib/driver_b200_single
export_rep() {
  ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/raw_$1.csv 2>/dev/null
  ncu -i $O/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_$1.csv.gz
  rm -f $O/prof_$1.ncu-rep
}
for k in k_sor_ring k_sched_tile k_collide k_prep; do
  OB_SCHED_TILE=8 ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_r02b_$k \
      $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_r02b_$k.log 2>&1
  export_rep r02b_$k
done
ls -la $O | tail -12
