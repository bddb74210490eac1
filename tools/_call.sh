for n in 8 4 2; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2950$n bench.py --gpus $n --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2p_final_cfg3_${n}gpu.json 2> gpurun_out/r2p_cfg3_${n}gpu.err
  tail -1 gpurun_out/r2p_final_cfg3_${n}gpu.json | cut -c1-230
done
python bench.py --steps 20 --warmup 5 --no-secondary --no-cpu-baseline > gpurun_out/r2p_final_cfg3_1gpu_same_box.json 2>/dev/null; tail -1 gpurun_out/r2p_final_cfg3_1gpu_same_box.json | cut -c1-230
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 200 --warmup 10 --no-secondary --no-cpu-baseline > gpurun_out/r2p_final_cfg3_8gpu_s200.json 2>/dev/null; tail -1 gpurun_out/r2p_final_cfg3_8gpu_s200.json | cut -c1-230
