#!/bin/sh
# Runs ON THE GPU BOX with N GPUs (gpurun --gpus N -- sh profiles/job_multi.sh TAG N): the default bench under torchrun
# (weak scaling over independent worlds), the split-world tests, and config 5 split over 1..N GPUs.
TAG=${1:-r01} ; N=${2:-2} ; O=gpurun_out ; mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -12 > $O/topo_${TAG}.txt
timeout 700 python -m pytest tests/test_split_world.py tests/test_multi_rank.py -m gpu -x -q -rs > $O/split_tests_${TAG}.log 2>&1; tail -6 $O/split_tests_${TAG}.log
python bench.py --steps 30 --warmup 3 --no-cpu > $O/bench_${TAG}_c2_n1.json 2> $O/bench_${TAG}_c2_n1.err
timeout 300 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu > $O/bench_${TAG}_c5_n1.json 2> $O/bench_${TAG}_c5_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 30 --warmup 3 > $O/bench_${TAG}_c2_n$n.json 2> $O/bench_${TAG}_c2_n$n.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --config 5 --gpus $n --steps 20 --warmup 3 --no-cpu > $O/bench_${TAG}_c5_n$n.json 2> $O/bench_${TAG}_c5_n$n.err
  tail -n 2 $O/bench_${TAG}_c2_n$n.err $O/bench_${TAG}_c5_n$n.err
done
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/bench_${TAG}_c*_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["n_gpus"], "%.3e" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], (r.get("phases_ms") or {}).get("sor"))
    except Exception as e:
        print(f, "ERR", e)
PY
