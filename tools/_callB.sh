set -x
python -m pytest tests -m gpu -q --tb=short -x 2>&1 | tail -15 > gpurun_out/b_gputests.log
(cd ab_old && python bench.py --workload dp --steps 6 --warmup 3 --no-cpu) > gpurun_out/b_dp1_old.json 2> gpurun_out/b_dp1_old.err
python bench.py --workload dp --steps 6 --warmup 3 --no-cpu > gpurun_out/b_dp1_new.json 2> gpurun_out/b_dp1_new.err
python bench.py --workload dp --dp-batch 1024 --steps 6 --warmup 3 --no-cpu > gpurun_out/b_dp1_new_b1024.json 2> gpurun_out/b_dp1_new_b1024.err
(cd ab_old && python bench.py --workload dp --dp-batch 1024 --steps 6 --warmup 3 --no-cpu) > gpurun_out/b_dp1_old_b1024.json 2> gpurun_out/b_dp1_old_b1024.err
cat gpurun_out/b_gputests.log
cut -c1-200 gpurun_out/b_dp1_old.json gpurun_out/b_dp1_new.json gpurun_out/b_dp1_old_b1024.json gpurun_out/b_dp1_new_b1024.json
