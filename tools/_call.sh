python -m pytest tests -m gpu -x -q -k "first_layer or vgg_bit_exact" > gpurun_out/r2j_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_tests.log
tail -3 gpurun_out/r2j_tests.log
for n in 296 1036 2072; do python tools/k5_probe.py $n 64 1 4; done 2>&1 | tee gpurun_out/r2j_k5.log
python tools/k5_probe.py 256 64 1 1 2>&1 | tee -a gpurun_out/r2j_k5.log
export QNNB_LIB=$PWD/quantizedneuralnetworks-keras-tensorflow_b200/libqnnb200_trace.so
export TRACE_WARPS=0,1,9,17 TRACE_TILES=9
python tools/k5_trace.py 1036 64 1 4 > gpurun_out/r2j_trace0.log 2>&1
