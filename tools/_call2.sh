set -x
python -m pytest tests -m gpu -q --tb=short -k "data_parallel" 2>&1 | tail -5 > gpurun_out/r02_2gpu_tests.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu > gpurun_out/r02_bench_2gpu.json 2> gpurun_out/r02_bench_2gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus 2 --workload dp --steps 6 --warmup 3 > gpurun_out/r02_bench_dp_2gpu.json 2> gpurun_out/r02_bench_dp_2gpu.err
cat gpurun_out/r02_2gpu_tests.log; cut -c1-300 gpurun_out/r02_bench_2gpu.json; cut -c1-300 gpurun_out/r02_bench_dp_2gpu.json
