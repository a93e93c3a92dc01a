#!/bin/bash
# developer loop: GPU suite + C2 bench line (no CPU baseline, no c5/module blocks) + batch-size sweep
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-quick}
timeout 1200 python -m pytest tests -m gpu -q -x --tb=short 2>&1 | grep -v "^E    +" | tail -15 > gpurun_out/${tag}_pytest.log
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c5 --no-module > gpurun_out/${tag}_bench.json 2> gpurun_out/${tag}_bench.err
timeout 300 python tools/gpu_bsweep.py 1 74 148 256 > gpurun_out/${tag}_bsweep.log 2>&1
tail -4 gpurun_out/${tag}_pytest.log; cat gpurun_out/${tag}_bsweep.log
python - "$tag" <<'PY'
import json,sys
d=json.loads(open(f"gpurun_out/{sys.argv[1]}_bench.json").read().strip().splitlines()[-1]); r=d["roofline"]
print("C2 step_ms=%.4f kernel_ms=%.4f frac=%.3f e2e_ms=%.3f" % (d["ms_per_step"], r["kernel_ms"], r["frac"], d["e2e"]["ms_per_step"]))
PY
