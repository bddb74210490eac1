timeout 900 python -m pytest tests -m gpu -x -q -k "first_layer or vgg_bit_exact or full_size or sm_share or intermediate" > gpurun_out/y1_tests.log 2>&1; echo "rc=$?" >> gpurun_out/y1_tests.log; tail -4 gpurun_out/y1_tests.log
for n in 296 1036 2072; do python tools/k5_probe.py $n 64 1 4 2>&1 | tail -1; done
python tools/k5_probe.py 4096 256 0 8 2>&1 | tail -1
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/y1_default.log 2>/dev/null; tail -1 gpurun_out/y1_default.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'], d['roofline']['per_kernel_ms'])"
python bench.py --no-secondary --no-cpu-baseline > gpurun_out/y1_s200.log 2>/dev/null; tail -1 gpurun_out/y1_s200.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(round(d['value']), d['ms_per_step'])"
