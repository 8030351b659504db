#!/bin/sh
# r02zz (GPU box): the whole GPU suite on the round's final tree + smoke()
O=gpurun_out ; mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r02zz_tests.log 2>&1
tail -4 $O/r02zz_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $O/r02zz_smoke.log 2>&1; tail -3 $O/r02zz_smoke.log
