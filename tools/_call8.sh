set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29622 bench.py --gpus 8 --workload table1 --steps 1 --warmup 0 > gpurun_out/r02_bench_table1_8gpu.json 2> gpurun_out/r02_bench_table1_8gpu.err
$T --master-port 29621 bench.py --gpus 8 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
cut -c1-600 gpurun_out/r02_bench_table1_8gpu.json; cut -c1-300 gpurun_out/r02_bench_8gpu.json; tail -3 gpurun_out/r02_bench_table1_8gpu.err
