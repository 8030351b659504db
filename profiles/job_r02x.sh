#!/bin/sh
# r02x (GPU box): k_sched_tile with the rows-per-level count taken out of the serial level chain: parity subset + configs[1], [3]
O=gpurun_out ; mkdir -p $O
timeout 400 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "sampled or full_batch or tile_width or (golden_reference_traces and (stack32 or ragdoll or hinges or crashwall or tower64 or mixed_maxc4) and single and not dropin)" > $O/r02x_tests.log 2>&1; tail -2 $O/r02x_tests.log
for t in c2:2 c4:4; do
  tag=${t%%:*}; cfg=${t#*:}
  python bench.py --config $cfg --steps 30 --warmup 3 --no-cpu --no-other > $O/r02x_$tag.json 2> $O/r02x_$tag.err
  python - "$O/r02x_$tag.json" "$tag" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
k=d["roofline"]["kernels"]
print(sys.argv[2], "ms/step %.3f"%d["ms_per_step"], " ".join("%s=%.3f"%(n,v["ms"]) for n,v in k.items()), "e2e %.3g"%d["e2e"]["value"])
PY
done
