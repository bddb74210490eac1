timeout 900 python -m pytest tests/test_gpu_models.py -q -x -k "float or trained" > gpurun_out/t9_float.log 2>&1; echo "rc=$?" >> gpurun_out/t9_float.log
tail -15 gpurun_out/t9_float.log
