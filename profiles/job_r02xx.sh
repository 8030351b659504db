#!/bin/sh
# r02xx (GPU box): ncu --set full of the two kernels that changed after the closing run (k_sched_tile, k_post on its own tile), configs[1]
O=gpurun_out ; mkdir -p $O
D=ode-0.12_b200/lib/driver_b200_single
cap() {
  timeout 120 ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02xx_$1 \
      $D --scene $4 --worlds $5 $6 --steps 4 --settle 150 --mode batch --time > $O/ncu_r02xx_$1.log 2>&1
  ncu -i $O/prof_r02xx_$1.ncu-rep --page raw --csv > $O/raw_r02xx_$1.csv 2>/dev/null
  rm -f $O/prof_r02xx_$1.ncu-rep
}
cap k_sched_tile k_sched_tile 152 stack32 4096 "--contacts-cap 192"
cap k_post k_post 152 stack32 4096 "--contacts-cap 192"
