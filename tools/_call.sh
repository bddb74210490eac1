timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/s24_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s24_tests.log
tail -3 gpurun_out/s24_tests.log
for rep in 1 2; do
timeout 300 python bench.py --workload cfg3 --streams 1 --steps 200 --warmup 10 --no-cpu-baseline > gpurun_out/s24_cfg3_$rep.log 2>&1
done
grep -h value gpurun_out/s24_cfg*.log | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l); r=d['roofline']; print(d['config']['workload'][:5], d['config']['streams'], round(d['value']), round(d['e2e']['value']), d['config']['parity_vs_exact_oracle'], {k:round(v*1e3,1) for k,v in r['per_kernel_ms'].items()})
"
