#!/bin/sh
# r02v (GPU box): k_sor_lane with lambda back in global memory + L2 prefetch of the row records six rows ahead: parity tests, configs[2]
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "sweep_matches_golden or sampled or full_batch" > $O/r02v_tests.log 2>&1
tail -3 $O/r02v_tests.log
for t in c3:X=1 c3_nolane:OB_SOR_LANE=0; do
  tag=${t%%:*}; envs=${t#*:}
  env $envs python bench.py --config 3 --steps 30 --warmup 3 --no-cpu --no-other > $O/r02v_$tag.json 2> $O/r02v_$tag.err
  python - "$O/r02v_$tag.json" "$tag" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
PY
done
