timeout 900 python -m pytest tests -m gpu -x -q -k "first_layer or vgg_bit_exact or full_size or intermediate" > gpurun_out/y2_tests.log 2>&1; echo "rc=$?" >> gpurun_out/y2_tests.log; tail -4 gpurun_out/y2_tests.log
for n in 296 1036 2072; do python tools/k5_probe.py $n 64 1 4 2>&1 | tail -1; done
python tools/k5_probe.py 4096 256 0 8 2>&1 | tail -1
