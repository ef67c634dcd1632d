set -x
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-strong --no-verify --cpu-budget 0 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
export BP_UNITS_PER_WARP=1
for c in 5,1,2 5,2,2 6,1,2 6,2,2; do
  n=$(echo $c | tr , _)
  python tools/profile_rollout.py --lcls $c --n-problems 150 --n-sim 960 > gpurun_out/final_plain_$n.json 2>&1 && \
  ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__thread_inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fs_rollout_lanes --csv --log-file gpurun_out/ncu_final_inst_$n.csv python tools/profile_rollout.py --lcls $c --n-problems 150 --n-sim 960 > gpurun_out/ncu_inst_$n.log 2>&1; echo "inst $c exit $?"
done
ncu --set full --clock-control none --import-source on -k regex:fs_rollout_lanes -s 1 -c 1 -o gpurun_out/prof_final_5_1_2 python tools/profile_rollout.py --lcls 5,1,2 --n-problems 505 --n-sim 1000 > gpurun_out/ncu_full.log 2>&1; echo "full exit $?"
ls -la gpurun_out | tail -20
