timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/s39_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s39_tests.log
tail -3 gpurun_out/s39_tests.log
timeout 300 python bench.py --workload cfg3 --streams 1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/s39_cfg3.log 2>&1
timeout 300 python bench.py --workload cfg3 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/s39_cfg3_4s.log 2>&1
timeout 300 python bench.py --workload cfg2 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/s39_cfg2.log 2>&1
grep -h '"value"' gpurun_out/s39_cfg*.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:5], d['config']['streams'], round(d['value']), d['config']['parity_vs_exact_oracle'], {k:round(v*1e3,1) for k,v in r['per_kernel_ms'].items()})
"
