python bench.py --workload cfg1 --no-cpu-baseline --no-secondary > gpurun_out/u2_cfg1.log 2>gpurun_out/u2_cfg1.err; tail -1 gpurun_out/u2_cfg1.log | cut -c1-900
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-secondary > gpurun_out/u2_cfg1_s20.log 2>gpurun_out/u2_cfg1_s20.err; tail -1 gpurun_out/u2_cfg1_s20.log | cut -c1-300
python bench.py --workload cfg1 --streams 2 --no-cpu-baseline --no-secondary > gpurun_out/u2_cfg1_ns2.log 2>/dev/null; tail -1 gpurun_out/u2_cfg1_ns2.log | cut -c1-300
python bench.py --workload cfg1 --streams 8 --no-cpu-baseline --no-secondary > gpurun_out/u2_cfg1_ns8.log 2>/dev/null; tail -1 gpurun_out/u2_cfg1_ns8.log | cut -c1-300
export QNNB_LIB=$PWD/quantizedneuralnetworks-keras-tensorflow_b200/libqnnb200_trace.so
TRACE_WARPS=0,2,16 python tools/net_trace.py 100 cfg1 > gpurun_out/u2_trace_n100.log 2>&1
