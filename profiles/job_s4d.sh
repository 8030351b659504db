#!/bin/sh
# GPU box job (2 GPUs): new joint scenes parity, split tests, config 5 at N=1,2 for two world sizes, default bench
O=gpurun_out; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pistons or pus" > $O/joint_tests_s4d.log 2>&1; tail -3 $O/joint_tests_s4d.log
timeout 600 python -m pytest tests/test_split_world.py -m gpu -x -q -rs > $O/split_tests_s4d.log 2>&1; tail -6 $O/split_tests_s4d.log
python bench.py --steps 30 --warmup 3 > $O/bench_s4d_c2.json 2> $O/bench_s4d_c2.err
for sc in pile_100x100x20 pile_200x200x20; do
  timeout 400 python bench.py --config 5 --scene $sc --steps 20 --warmup 3 --no-cpu > $O/bench_s4d_c5_${sc}_n1.json 2> $O/bench_s4d_c5_${sc}_n1.err
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --config 5 --scene $sc --gpus 2 --steps 20 --warmup 3 --no-cpu > $O/bench_s4d_c5_${sc}_n2.json 2> $O/bench_s4d_c5_${sc}_n2.err
  tail -2 $O/bench_s4d_c5_${sc}_n1.err $O/bench_s4d_c5_${sc}_n2.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_s4d_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["n_gpus"], "%.3e" % d["value"], "%.3f ms" % d["ms_per_step"], r.get("phases_ms") or {k: round(v["ms"], 3) for k, v in r["kernels"].items()})
    except Exception as e:
        print(f, "ERR", e)
PY
