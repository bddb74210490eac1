python -m pytest tests -m gpu -x -q > gpurun_out/t1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t1_tests.log
tail -5 gpurun_out/t1_tests.log
python bench.py --streams 1 --no-secondary --no-cpu-baseline > gpurun_out/t1_cfg3_s1.log 2>gpurun_out/t1_cfg3_s1.err; tail -1 gpurun_out/t1_cfg3_s1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['roofline']['per_kernel_ms'])"
python bench.py > gpurun_out/t1_default.log 2>gpurun_out/t1_default.err; tail -1 gpurun_out/t1_default.log | cut -c1-600
