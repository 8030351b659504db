#!/bin/sh
# r02n (GPU box with N GPUs: gpurun --gpus N -- sh profiles/job_r02n.sh N): large-world front-end split with bulk push kernels; crashwall (two policy rows) on the GPU
# narrowphase by SAP-sorted position, peer stores + two flag barriers per step): loop-back and one-device-per-rank tests, nested-space
# scenes in dDOUBLE, configs[4] at N = 1 .. N on the same box (both split modes), the default bench at N = 1 and N.
N=${1:-2} ; O=gpurun_out ; mkdir -p $O
nvidia-smi topo -m 2>/dev/null | head -12 > $O/topo_r02n.txt
timeout 900 python -m pytest tests/test_split_world.py tests/test_large_world.py tests/test_multi_rank.py -m gpu -q -rs > $O/r02n_split_tests.log 2>&1; tail -8 $O/r02n_split_tests.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "crashwall or nested" > $O/r02n_nested.log 2>&1; tail -2 $O/r02n_nested.log
timeout 300 python bench.py --config 5 --steps 20 --warmup 3 --no-cpu > $O/r02n_c5_n1.json 2> $O/r02n_c5_n1.err
for n in 2 4 8; do
  [ $n -le $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --config 5 --gpus $n --steps 20 --warmup 3 --no-cpu > $O/r02n_c5_n$n.json 2> $O/r02n_c5_n$n.err
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --config 5 --scene pile_200x200x20 --gpus $n --steps 10 --warmup 3 --no-cpu > $O/r02n_c5big_n$n.json 2> $O/r02n_c5big_n$n.err
done
timeout 300 python bench.py --config 5 --scene pile_200x200x20 --steps 10 --warmup 3 --no-cpu > $O/r02n_c5big_n1.json 2> $O/r02n_c5big_n1.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29540 bench.py --gpus $N --steps 20 --warmup 5 > $O/r02n_c2_n$N.json 2> $O/r02n_c2_n$N.err
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r02n_c*_n*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        r = d["roofline"]
        print(f, d["n_gpus"], "%.3e" % d["value"], "%.3f ms" % d["ms_per_step"], "e2e %.3e" % d["e2e"]["value"], {k: round(v, 3) for k, v in (r.get("phases_ms") or {}).items()})
        for k, v in (d.get("other_configs") or {}).items():
            print("   ", k, v and {x: v.get(x) for x in ("ms_per_step", "value", "e2e")})
    except Exception as e:
        print(f, "ERR", e)
PY
