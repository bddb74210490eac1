timeout 900 python -m pytest tests/test_gpu_fused_net.py tests/test_gpu_models.py -q -x > gpurun_out/t8_tests.log 2>&1; echo "rc=$?" >> gpurun_out/t8_tests.log
tail -4 gpurun_out/t8_tests.log
python bench.py --workload cfg1 --no-secondary > gpurun_out/t8_cfg1.log 2>gpurun_out/t8_cfg1.err; tail -1 gpurun_out/t8_cfg1.log | cut -c1-2500
