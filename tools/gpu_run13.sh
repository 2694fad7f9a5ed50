python -m pytest tests/test_cluster_gpu.py -m gpu -x -q 2>&1 | tail -3
for g in 0 1; do
  if [ $g = 1 ]; then export STRGPU_NO_GRAPH=1; fi
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli --parity-sample 0 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_graph=$g value',d['value'],'ms',d['ms_per_step'],'cluster_ms',d['cluster']['ms_per_step'])"
  python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-cli --parity-sample 0 --reads-per-gpu 72000000 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('no_graph=$g small: ms',d['ms_per_step'],'cluster_ms',d['cluster']['ms_per_step'], d['cluster']['treads_per_gpu'])"
done
