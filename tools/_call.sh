timeout 600 python -m pytest tests/test_gpu_fused_net.py -x -q > gpurun_out/t2_fused.log 2>&1; echo "rc=$?" >> gpurun_out/t2_fused.log
tail -30 gpurun_out/t2_fused.log
