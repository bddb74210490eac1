#!/bin/bash
# Round-2 evidence run (one GPU): bench lines, ncu launch lists and full-set captures of one forward per workload.
# usage (GPU box): bash tools/profile_r2.sh   -> gpurun_out/r2p_*
O=gpurun_out
B="--no-cpu-baseline --no-secondary"
python bench.py --steps 20 --warmup 5 > $O/r2p_final_cfg3_default.json 2> $O/r2p_cfg3_default.err
python bench.py --streams 1 $B > $O/r2p_final_cfg3_s1.json 2>/dev/null
for w in cfg1 cfg2 cfg4 cfg5 cfg5t; do python bench.py --workload $w --no-secondary > $O/r2p_final_$w.json 2>/dev/null; done
python bench.py --workload cfg2 --batch 1024 $B > $O/r2p_final_cfg2_b1024.json 2>/dev/null
python bench.py --workload cfg1 --batch 4096 $B > $O/r2p_final_cfg1_b4096.json 2>/dev/null
python bench.py --impl reference --steps 3 --warmup 1 > $O/r2p_final_ref.json 2>/dev/null
K='regex:conv3x3|conv_generic|dense_|vgg_fused'
for spec in cfg3:4 cfg1:1 cfg4:10 cfg5:12 cfg2:4; do
  w=${spec%%:*}; c=${spec##*:}
  A="--workload $w --steps 2 --warmup 3 --graphs 0 --streams 1 $B"
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2p_launches_$w.csv python bench.py $A > $O/r2p_ncu_launch_$w.log 2>&1
  ncu --set full --clock-control none --import-source on -k "$K" -c $c -f -o /tmp/r2p_prof_$w python bench.py $A > $O/r2p_ncu_full_$w.log 2>&1
  python tools/ncu_summary.py /tmp/r2p_prof_$w.ncu-rep $O/r2p_ncu_$w.csv >> $O/r2p_ncu_full_$w.log 2>&1
done
cp /tmp/r2p_prof_cfg1.ncu-rep /tmp/r2p_prof_cfg3.ncu-rep $O/   # the two small captures travel back (source view); gpurun_out is capped at 64 MiB
ls -la $O/r2p_* | awk '{print $5, $9}'
