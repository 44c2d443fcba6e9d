set -x
T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
$T --master-port 29621 bench.py --gpus 8 --steps 6 --warmup 3 > gpurun_out/r02_bench_8gpu.json 2> gpurun_out/r02_bench_8gpu.err
$T --master-port 29622 bench.py --gpus 8 --workload table1 --steps 1 --warmup 0 > gpurun_out/r02_bench_table1_8gpu.json 2> gpurun_out/r02_bench_table1_8gpu.err
( time $T --master-port 29623 mr_gan.py --tables 1 --seed 0 --synthetic --precision f16 ) > gpurun_out/r02_table1_8gpu_cli.log 2>&1
cut -c1-400 gpurun_out/r02_bench_8gpu.json; cut -c1-600 gpurun_out/r02_bench_table1_8gpu.json; tail -8 gpurun_out/r02_table1_8gpu_cli.log
