timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo bench rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json'))
for k in ('value','ms_per_step','gpu_launches','roofline','cluster','cpu_baseline','parity_sample','cli','clocks'): print(k, d.get(k))
print('e2e', d['e2e'])
"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null; head -c 400 gpurun_out/bench_ref.json
