cd "$GRAFT_REPO_ROOT" || exit 1
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "host_session" 2>&1 | tail -2
for sl in 8 16 32 64; do python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-c5 --no-module --slices $sl 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); e=d['e2e']; print('slices', $sl, 'e2e_ms', round(e['ms_per_step'],4), 'value', round(e['value']/1e9,4), 'h2d MB', round(e['h2d_bytes_per_step']/1e6,2), 'launches', e['launches_per_step'])"; done
