#!/bin/sh
# r02w (GPU box): ncu --set full of the collide kernels of configs[2] on the final tree (lane use after the lock-step BVH loops)
O=gpurun_out ; mkdir -p $O
D=ode-0.12_b200/lib/driver_b200_single
cap() {
  timeout 150 ncu --set full --clock-control none -k regex:$2 -s $3 -c 1 -f -o $O/prof_r02w_$1 \
      $D --scene $4 --worlds $5 $6 --steps 4 --settle 60 --mode batch --time > $O/ncu_r02w_$1.log 2>&1
  ncu -i $O/prof_r02w_$1.ncu-rep --page raw --csv > $O/raw_r02w_$1.csv 2>/dev/null
  rm -f $O/prof_r02w_$1.ncu-rep
}
cap c3_k_narrow k_narrow 62 buggy_terrain256 65536 "--contacts-cap 48"
cap c3_k_broad_tile k_broad_tile 62 buggy_terrain256 65536 "--contacts-cap 48"
ls -la $O | grep r02w
