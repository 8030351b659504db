#!/bin/sh
# A/B on one box: current library vs ode-0.12_b200/lib_base/libode_b200_single.so (a build of an earlier commit,
# staged by hand, not tracked), interleaved twice, configs 2 and 4, to price a change of the collide kernels
O=gpurun_out; mkdir -p $O
L=ode-0.12_b200/lib/libode_b200_single.so
S=ode-0.12_b200/lib/libob_scenes_single.so
cp $L /tmp/cur.so; cp $S /tmp/cur_scenes.so
for rep in 1 2; do
  for v in cur base; do
    if [ $v = base ]; then cp ode-0.12_b200/lib_base/libode_b200_single.so $L; cp ode-0.12_b200/lib_base/libob_scenes_single.so $S; else cp /tmp/cur.so $L; cp /tmp/cur_scenes.so $S; fi
    for c in 2 4; do
      timeout 100 python bench.py --config $c --steps 30 --warmup 3 --no-cpu > $O/ab_${v}_${rep}_c$c.json 2> $O/ab_${v}_${rep}_c$c.err
      python -c "
import json
d=json.loads(open('$O/ab_${v}_${rep}_c$c.json').read().strip().splitlines()[-1]); print('$v', $rep, $c, round(d['ms_per_step'],4), {k: round(x['ms'],4) for k,x in d['roofline']['kernels'].items()})" || tail -3 $O/ab_${v}_${rep}_c$c.err
    done
  done
done
cp /tmp/cur.so $L; cp /tmp/cur_scenes.so $S
