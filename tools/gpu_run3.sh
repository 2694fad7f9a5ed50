python -m pytest tests/test_scan_gpu.py -m gpu -x -q 2>&1 | tail -3
python tools/profile_scan.py --sweep 2>&1 | grep us_per_call
for s in 4; do python bench.py --steps 3 --warmup 3 --no-cpu-baseline --streams $s 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('streams',$s,'value',d['value'],'ms',d['ms_per_step'],'frac',d['roofline']['frac'],'cluster_ms',d['cluster']['ms_per_step'])"; done
