# Multi-GPU measurement batch: bash tools/measure_multi.sh N   (through gpurun --gpus N)
N=$1
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
$T --master-port 29521 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/r2_bench_n$N.json 2>gpurun_out/r2_bench_n$N.err
$T --master-port 29522 bench.py --gpus $N --workload c3 --steps 4 --warmup 1 > gpurun_out/r2_bench_n${N}_c3.json 2>gpurun_out/r2_bench_n${N}_c3.err
[ -n "$SKIP_MULTI" ] || $T --master-port 29523 tools/run_multi.py > gpurun_out/r2_multi_n$N.json 2>gpurun_out/r2_multi_n$N.err
tail -c 300 gpurun_out/r2_multi_n$N.err
