export QNNB_LIB=$PWD/quantizedneuralnetworks-keras-tensorflow_b200/libqnnb200_trace.so
export TRACE_WARPS=0,1,5,9,13,17 TRACE_TILES=7
python tools/k5_trace.py 1036 64 1 4 > gpurun_out/w2_trace.log 2>&1
for e in 0 1 9 15 13; do QNNB_K5_EXP=$e python tools/k5_probe.py 2072 64 1 4 2>&1 | tail -1; done
