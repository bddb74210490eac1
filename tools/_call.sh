set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/h_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/h_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1
timeout 300 python bench.py > gpurun_out/h_cfg3_default.log 2>&1
timeout 300 python bench.py --streams 1 --no-cpu-baseline > gpurun_out/h_cfg3_s1.log 2>&1
timeout 300 python bench.py --workload cfg4 --steps 20 --warmup 3 > gpurun_out/h_cfg4.log 2>&1
timeout 300 python bench.py --workload cfg2 > gpurun_out/h_cfg2.log 2>&1
timeout 300 python bench.py --workload cfg2 --batch 1024 --no-cpu-baseline > gpurun_out/h_cfg2_b1024.log 2>&1
timeout 300 python bench.py --workload cfg1 > gpurun_out/h_cfg1.log 2>&1
timeout 300 python bench.py --workload cfg5 --steps 30 --warmup 3 > gpurun_out/h_cfg5.log 2>&1
timeout 300 python bench.py --workload cfg5t --steps 30 --warmup 3 --no-cpu-baseline > gpurun_out/h_cfg5t.log 2>&1
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/h_ref.log 2>&1
mkdir -p /tmp/ncu
B="python bench.py --workload cfg3 --steps 2 --warmup 3 --graphs 0 --streams 1 --no-cpu-baseline"
$B > gpurun_out/h_plain_cfg3.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/h_launches_cfg3.csv $B > /tmp/ncu/l.log 2>&1
$B > /tmp/ncu/plain.log 2>&1 && ncu --set full --clock-control none -k regex:'conv|dense' -s 12 -c 4 -o /tmp/ncu/prof_cfg3 -f $B > /tmp/ncu/full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/prof_cfg3.ncu-rep gpurun_out/h_ncu_cfg3.csv
rm -rf /tmp/ncu
set +x
tail -2 gpurun_out/h_tests.log; cat gpurun_out/h_smoke.log | tail -2
grep -h '"value"' gpurun_out/h_*.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l)
    if d.get('impl')=='reference': print('REF', round(d['value']), d['cpu_baseline']['cores']); continue
    r=d['roofline']; print(d['config']['workload'][:34], 's=%d'%d['config']['streams'], round(d['value']), 'e2e', round(d['e2e']['value']), d['config']['parity_vs_exact_oracle'], d['clocks']['reasons'], r['kernel'][:30], r['bound'], round(r['achieved'],1), round(r['frac'],3), round(r['share_of_step'],2), 'cpu', round(d['cpu_baseline']['value']) if 'cpu_baseline' in d else None, {k:round(v*1e3,1) for k,v in list(r['per_kernel_ms'].items())[:10]})
"
