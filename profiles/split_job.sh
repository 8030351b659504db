#!/bin/sh
# Runs ON THE GPU BOX (gpurun [--gpus N] -- sh profiles/split_job.sh TAG N): split-world tests, then bench.py --config 5 at 1..N GPUs.
TAG=${1:-r01s}
N=${2:-1}
O=gpurun_out
mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -12 > $O/topo_${TAG}.txt
timeout 700 python -m pytest tests/test_split_world.py -m gpu -x -q -rs > $O/split_tests_${TAG}.log 2>&1
tail -12 $O/split_tests_${TAG}.log
timeout 300 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu > $O/bench_${TAG}_c5_n1.json 2> $O/bench_${TAG}_c5_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --config 5 --gpus $n --steps 20 --warmup 3 --no-cpu > $O/bench_${TAG}_c5_n$n.json 2> $O/bench_${TAG}_c5_n$n.err
  tail -3 $O/bench_${TAG}_c5_n$n.err
done
python - <<PY
import json
for n in (1, 2, 4, 8):
    try:
        d = json.loads(open(f"$O/bench_${TAG}_c5_n{n}.json").read().strip().splitlines()[-1])
        print(n, d["value"], d["ms_per_step"], d["e2e"]["value"], d["roofline"]["phases_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
