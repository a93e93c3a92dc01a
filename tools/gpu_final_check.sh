cd "$GRAFT_REPO_ROOT" || exit 1
python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print({k:d[k] for k in ('value','ms_per_step','n_gpus','gpu_launches')}, d['roofline']['frac'], d['roofline']['kernel_ms'], d['e2e']['value'], d['e2e']['ms_per_step'], d['e2e']['h2d_bytes_per_step'], d['c5']['ms_per_step'], d['cpu_baseline']['value'], d['clocks'])"
python bench.py --impl reference --steps 5 --warmup 2 2>/dev/null | cut -c1-300
