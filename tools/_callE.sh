set -x
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -5 > gpurun_out/e_gputests.log
python bench.py --workload dp --steps 6 --warmup 3 --no-cpu > gpurun_out/e_dp1.json 2> gpurun_out/e_dp1.err
python bench.py --workload dp --dp-batch 1024 --steps 6 --warmup 3 --no-cpu > gpurun_out/e_dp1_b1024.json 2> gpurun_out/e_dp1_b1024.err
python bench.py --no-cpu --no-modes > gpurun_out/e_bench.json 2> gpurun_out/e_bench.err
cat gpurun_out/e_gputests.log
cut -c1-200 gpurun_out/e_dp1.json gpurun_out/e_dp1_b1024.json gpurun_out/e_bench.json
