#!/bin/bash
# developer tool: the bench lines, launch list and ncu capture committed under profiles/ (tag = $1)
tag=$1
python bench.py > gpurun_out/${tag}_bench_c2.json 2> gpurun_out/${tag}_bench_c2.err
python bench.py --impl reference > gpurun_out/${tag}_ref.json 2>/dev/null
for w in C1 C3 C4; do python bench.py --workload $w --steps 50 --no-cpu-baseline > gpurun_out/${tag}_bench_$w.json 2>/dev/null; done
python bench.py --peaky --steps 50 --no-cpu-baseline > gpurun_out/${tag}_bench_c2_peaky.json 2>/dev/null
python bench.py --workload C5 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/${tag}_bench_c5_n1.json 2>/dev/null
python bench.py --steps 3 --warmup 3 --no-cpu-baseline > /dev/null 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/${tag}_launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 3 -c 1 -f -o gpurun_out/prof_${tag}_c2 \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:ctc_lin -s 3 -c 1 -f -o gpurun_out/prof_${tag}_c4 \
    python bench.py --workload C4 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/ncu3.log 2>&1
python tools/gpu_comparators.py --out gpurun_out/${tag}_comparators.json > gpurun_out/cmp.log 2>&1
for f in gpurun_out/${tag}_bench_*.json; do python - "$f" <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); r=d["roofline"]
print(sys.argv[1].split("/")[-1], "value=%.4g"%d["value"], "step_ms=%.4f"%d["ms_per_step"], "kernel_ms=%.4f"%r["kernel_ms"], "frac=%.3f"%r["frac"], "e2e_ms=%.3f"%d["e2e"]["ms_per_step"], "e2e=%.4g"%d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"))
PY
done
