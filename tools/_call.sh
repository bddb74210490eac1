python bench.py --steps 20 --warmup 5 > gpurun_out/v3_default.log 2>gpurun_out/v3_default.err; tail -1 gpurun_out/v3_default.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['e2e']['value'], d['roofline']['step_frac']); print({k:(round(v['value']), v['steps'], v['roofline']['frac'], v['roofline']['step_frac']) for k,v in d['secondary'].items()})"
tail -3 gpurun_out/v3_default.err
python bench.py --workload cfg2 --no-cpu-baseline --no-secondary > gpurun_out/v3_cfg2.log 2>/dev/null; tail -1 gpurun_out/v3_cfg2.log | cut -c1-250
python bench.py --workload cfg2 --batch 1024 --no-cpu-baseline --no-secondary > gpurun_out/v3_cfg2_b1024.log 2>/dev/null; tail -1 gpurun_out/v3_cfg2_b1024.log | cut -c1-250
