#!/bin/sh
# r02y (GPU box with 2 GPUs): k_post on its own tile width (parity subset + configs[1] at N = 1), then what the driver's SCALE step runs at
# N = 2: the reference arm and the default bench under torchrun (other_configs: configs[1] strong, [2], [3], [4] split), split-world tests
O=gpurun_out ; mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "golden" > $O/r02y_tests.log 2>&1; tail -2 $O/r02y_tests.log
timeout 600 python -m pytest tests/test_split_world.py tests/test_multi_rank.py -m gpu -q -rs > $O/r02y_split_tests.log 2>&1; tail -3 $O/r02y_split_tests.log
python bench.py --steps 30 --warmup 3 --no-cpu --no-other > $O/r02y_c2_n1.json 2> $O/r02y_c2_n1.err
OB_POST_TILE=8 python bench.py --steps 30 --warmup 3 --no-cpu --no-other > $O/r02y_c2_n1_post8.json 2> $O/r02y_c2_n1_post8.err
t0=$(date +%s)
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29561 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 > $O/r02y_ref_n2.json 2> $O/r02y_ref_n2.err
t1=$(date +%s); echo "reference arm N=2 wall $((t1-t0)) s"
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29562 bench.py --gpus 2 --steps 20 --warmup 5 > $O/r02y_c2_n2.json 2> $O/r02y_c2_n2.err
t2=$(date +%s); echo "b200 arm N=2 wall $((t2-t1)) s"
python - <<'PY'
import json
for f in ("r02y_c2_n1","r02y_c2_n1_post8","r02y_c2_n2","r02y_ref_n2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, d.get("n_gpus"), "ms/step %.3f"%d["ms_per_step"], "value %.3g"%d["value"], "e2e %.3g"%d["e2e"]["value"], {k:round(v["ms"],3) for k,v in d.get("roofline",{}).get("kernels",{}).items()})
        for k,v in (d.get("other_configs") or {}).items():
            print("   ", k, v and {x:v.get(x) for x in ("ms_per_step","value","e2e","nvlink_bytes_per_step_per_rank")})
    except Exception as e: print(f,"ERR",e)
PY
