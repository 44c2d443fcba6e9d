set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/r02_final_gputests.log
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-modes > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-cpu --no-modes > gpurun_out/r02_final_ncu_launches.log 2>&1
CMD="python tools/profile_epoch.py --folds 74 --precision f16"
$CMD > /dev/null 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_dw_adam_tc -s 0 -c 8 -o gpurun_out/r02_final_dw $CMD > gpurun_out/r02_final_ncu_dw.log 2>&1
cat gpurun_out/r02_final_gputests.log
tail -c 600 gpurun_out/r02_bench_1gpu.json
