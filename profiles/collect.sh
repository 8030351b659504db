#!/bin/sh
# Runs ON THE GPU BOX (gpurun -- sh profiles/collect.sh TAG): bench lines for every config, the ncu launch
# list of the default bench command and one `--set full` capture per dominant kernel.  Outputs go to
# gpurun_out/; profiles/summarize.py (run in the build container) turns them into the tracked summaries.
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
# ncu reports are ~20 MB each and gpurun brings back at most 64 MiB: every capture is turned into its raw-metrics CSV
# (+ the gzipped source page) on the box and the .ncu-rep is dropped
export_rep() {   # $1 = report stem
  ncu -i $O/prof_$1.ncu-rep --page raw --csv > $O/raw_$1.csv 2>/dev/null
  ncu -i $O/prof_$1.ncu-rep --page source --csv 2>/dev/null | gzip > $O/src_$1.csv.gz
  rm -f $O/prof_$1.ncu-rep
}
if [ -z "$PROF_ONLY" ]; then
python bench.py --steps 30 --warmup 3 > $O/bench_${TAG}_c2.json 2> $O/bench_${TAG}_c2.err
python bench.py --config 3 --steps 30 --warmup 3 > $O/bench_${TAG}_c3.json 2> $O/bench_${TAG}_c3.err
python bench.py --config 4 --steps 30 --warmup 3 > $O/bench_${TAG}_c4.json 2> $O/bench_${TAG}_c4.err
python bench.py --config 1 --steps 200 --warmup 5 > $O/bench_${TAG}_c1.json 2> $O/bench_${TAG}_c1.err
python bench.py --config 5 --steps 20 --warmup 3 > $O/bench_${TAG}_c5.json 2> $O/bench_${TAG}_c5.err
python bench.py --impl reference --steps 30 --warmup 3 > $O/bench_${TAG}_ref.json 2> $O/bench_${TAG}_ref.err
fi
# launch list of the same command as the default bench (numbers printed under ncu are never bench values)
ncu --metrics gpu__time_duration.sum --clock-control none -s 1600 -c 400 --csv --log-file $O/launches_${TAG}_c2.csv \
    python bench.py --steps 30 --warmup 3 > $O/ncu_${TAG}_c2.log 2>&1
D=ode-0.12_b200/lib/driver_b200_single
for k in k_sor k_collide k_prep k_sched k_post; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 305 -c 1 -f -o $O/prof_${TAG}_$k \
      $D --scene stack32 --worlds 4096 --contacts-cap 192 --steps 10 --settle 300 --mode batch --time > $O/ncu_${TAG}_$k.log 2>&1
  export_rep ${TAG}_$k
done
ncu --set full --clock-control none --import-source on -k regex:k_lw_sor_all -s 300 -c 1 -f -o $O/prof_${TAG}_k_lw_sor_all \
    $D --scene pile_100x100x20 --steps 3 --settle 300 --mode batch --time > $O/ncu_${TAG}_k_lw_sor_all.log 2>&1
export_rep ${TAG}_k_lw_sor_all
# config 3 (buggies on the shared terrain mesh): its dominant kernels
for k in k_collide k_sor; do
  ncu --set full --clock-control none --import-source on -k regex:$k -s 203 -c 1 -f -o $O/prof_${TAG}_c3_$k \
      $D --scene buggy_terrain256 --worlds 65536 --contacts-cap 48 --steps 6 --settle 200 --mode batch --time > $O/ncu_${TAG}_c3_$k.log 2>&1
  export_rep ${TAG}_c3_$k
done
ls -la $O | tail -30
