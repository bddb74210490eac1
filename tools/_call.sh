mkdir -p /tmp/ncu
B="python bench.py --workload cfg5 --steps 2 --warmup 3 --graphs 0 --streams 1 --no-cpu-baseline"
$B > gpurun_out/y_plain_cfg5.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/y_launches_cfg5.csv $B > /tmp/ncu/l.log 2>&1
$B > /tmp/ncu/plain.log 2>&1 && ncu --set full --clock-control none -k regex:'conv|dense' -s 192 -c 12 -o /tmp/ncu/prof_cfg5 -f $B > /tmp/ncu/full.log 2>&1
python tools/ncu_summary.py /tmp/ncu/prof_cfg5.ncu-rep gpurun_out/y_ncu_cfg5.csv
rm -rf /tmp/ncu
ls -la gpurun_out/y_*
