# The round's 1-GPU measurement batch (run through gpurun; outputs land in gpurun_out/, summaries are copied to profiles/).
set -x
python bench.py --impl reference > gpurun_out/r2_bench_reference_arm.json 2>gpurun_out/r2_bench_reference_arm.err
python bench.py > gpurun_out/r2_bench_n1.json 2>gpurun_out/r2_bench_n1.err
python bench.py --workload c2_paper8 --steps 10 > gpurun_out/r2_bench_n1_paper8.json 2>/dev/null
python bench.py --workload c3 --steps 3 --warmup 1 > gpurun_out/r2_bench_n1_c3.json 2>/dev/null
python tools/run_configs.py > gpurun_out/r2_configs.md 2>gpurun_out/r2_configs.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 3 > gpurun_out/ncu_launch.log 2>&1
ncu --set full --clock-control none -c 16 -o gpurun_out/r2_full python tools/quick_stage0.py 8 ncu > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
ls -la gpurun_out
