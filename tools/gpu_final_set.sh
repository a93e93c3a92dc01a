#!/bin/bash
# developer tool: the last profile set of a round on the committed tree (tag = $1): GPU suite, default bench line + reference
# arm, one line per secondary workload, launch list of the bench command, one ncu --set full capture of the R177 kernel
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
tag=${1:-r02_final}
timeout 1500 python -m pytest tests -m gpu -q -rs --tb=short 2>&1 | grep -v "^E    +" | tail -n 12 > gpurun_out/${tag}_pytest_gpu.log
python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --impl reference > gpurun_out/${tag}_ref.json 2>/dev/null
for w in C1 C3 C4 R177; do python bench.py --workload $w --steps 50 --no-cpu-baseline --no-c5 --no-module > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/ncu1.log 2>&1
python bench.py --workload R177 --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > /dev/null 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 3 -c 1 -f -o gpurun_out/prof_${tag}_R177 \
    python bench.py --workload R177 --steps 3 --warmup 3 --no-cpu-baseline --no-c5 --no-module > gpurun_out/ncu2.log 2>&1
tail -n 3 gpurun_out/${tag}_pytest_gpu.log
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1].split("/")[-1], "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%r["kernel_ms"], "frac=%.3f"%r["frac"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], "e2e=%.4g"%d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"))
PY
done
