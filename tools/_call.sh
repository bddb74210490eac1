timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/g_tests.log
tail -3 gpurun_out/g_tests.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/g_cfg3_default.log 2>&1
timeout 300 python bench.py --streams 1 --no-cpu-baseline > gpurun_out/g_cfg3_s1.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/g_cfg4.log 2>&1
timeout 300 python bench.py --workload cfg2 > gpurun_out/g_cfg2.log 2>&1
timeout 300 python bench.py --workload cfg2 --batch 1024 --no-cpu-baseline > gpurun_out/g_cfg2_b1024.log 2>&1
timeout 300 python bench.py --workload cfg1 > gpurun_out/g_cfg1.log 2>&1
timeout 300 python bench.py --workload cfg5 --steps 30 --warmup 3 > gpurun_out/g_cfg5.log 2>&1
timeout 300 python bench.py --workload cfg5t --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/g_cfg5t.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/g_ref.log 2>&1
grep -h '"value"' gpurun_out/g_*.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('impl')=='reference': print('REF', round(d['value']), d['cpu_baseline']['cores']); continue
    r=d['roofline']; print(d['config']['workload'][:34], 's=%d'%d['config']['streams'], round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['parity_vs_exact_oracle'], d['clocks']['reasons'], r['kernel'], r['bound'], round(r['frac'],3), 'traffic', r['traffic'], 'cpu', round(d['cpu_baseline']['value']) if 'cpu_baseline' in d else None)
"
