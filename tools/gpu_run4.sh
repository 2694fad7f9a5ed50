python -m pytest tests/test_cluster_gpu.py tests/test_cli_gpu.py -m gpu -x -q 2>&1 | tail -8
python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams 4 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'cluster',d['cluster'])"
