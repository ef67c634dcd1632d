# Final evidence of a round: GPU tests, smoke, the default bench line, the reference arm, other
# workloads, the ncu launch list of the bench command.  Run under gpurun from the repository root.
set -x
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/final_tests.log 2>&1; echo "pytest exit $?"; tail -2 gpurun_out/final_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; echo "smoke exit $?"; tail -1 gpurun_out/final_smoke.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_final_reference.json 2> gpurun_out/bench_final_reference.err; echo "bench reference exit $?"
python bench.py --set b --strategy avg_depth --cpu-budget 0 --no-extras > gpurun_out/bench_final_b.json 2> gpurun_out/bench_final_b.err; echo "bench b exit $?"
python bench.py --n-sim 100 --cpu-budget 0 --no-extras > gpurun_out/bench_final_n100.json 2>&1; echo "bench n100 exit $?"
python bench.py --n-sim 5000 --steps 5 --cpu-budget 0 --no-extras > gpurun_out/bench_final_n5000.json 2>&1; echo "bench n5000 exit $?"
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_final.csv python bench.py --steps 2 --warmup 3 --no-e2e --no-extras --no-verify --cpu-budget 0 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches exit $?"
python tools/solve_rate_table.py > gpurun_out/solve_rates.json 2> gpurun_out/solve_rates.err; echo "solve rates exit $?"
