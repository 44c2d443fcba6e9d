# One-GPU evidence run of a round: GPU tests, smoke, bench lines, launch list, ncu captures (summaries extracted on the box).
set -x
python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -5 > gpurun_out/r02_gputests_1gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1
python bench.py > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
python bench.py --impl reference > gpurun_out/r02_bench_reference_arm.json 2> gpurun_out/r02_bench_ref.err
python bench.py --workload dp --steps 6 --warmup 3 --no-cpu > gpurun_out/r02_bench_dp_1gpu.json 2> gpurun_out/r02_bench_dp_1gpu.err
CMD="python tools/profile_epoch.py --folds 74 --precision f16"
KR='regex:k_gemm_tc|k_dw_adam_tc|k_prep|k_adam|k_bn|k_fm|k_loss|k_argmax|k_epoch'
$CMD > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -k "$KR" -c 400 --csv --log-file gpurun_out/r02_f16_launches_74folds.csv $CMD > /dev/null 2>&1
export MRGAN_CHAINS=1
ncu --set full --clock-control none --import-source on -k regex:k_dw_adam_tc -c 3 -o /tmp/dw $CMD > /dev/null 2>&1
python tools/update_traffic.py /tmp/dw.ncu-rep f16 74 1200 profiles/r02_f16_dw1_ncu_full.txt > gpurun_out/r02_traffic.log 2>&1
cp profiles/ncu_traffic.json gpurun_out/ncu_traffic.json
L=$(python -c "import json; print(json.load(open('profiles/ncu_traffic.json'))['dw1/f16']['launch_in_report'])")
python profiles/extract_ncu.py /tmp/dw.ncu-rep $L > gpurun_out/r02_f16_dw1_ncu_full.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -c 16 -o /tmp/g $CMD > /dev/null 2>&1
ncu -i /tmp/g.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; ix={k:i for i,k in enumerate(h)}
fwd=dx=None
for n,r in enumerate(rows[2:]):
    name, grid = r[ix['Kernel Name']], r[ix['Grid Size']].replace(' ','')
    print('#', n, name[:48], grid, r[ix['gpu__time_duration.sum']], file=sys.stderr)
    if grid == '(4,1,74)' and 'k_gemm_tc<1, 0' in name and fwd is None: fwd = n
    if grid == '(4,1,74)' and 'k_gemm_tc<0, 0' in name and dx is None: dx = n
print(fwd if fwd is not None else -1, dx if dx is not None else -1)
" > gpurun_out/r02_gemm_idx.txt 2> gpurun_out/r02_gemm_launch_index.txt
read FWD DX < gpurun_out/r02_gemm_idx.txt
[ "$FWD" -ge 0 ] && python profiles/extract_ncu.py /tmp/g.ncu-rep $FWD > gpurun_out/r02_f16_fwd1_ncu_full.txt 2>&1
[ "$DX" -ge 0 ] && python profiles/extract_ncu.py /tmp/g.ncu-rep $DX > gpurun_out/r02_f16_dx2_ncu_full.txt 2>&1
unset MRGAN_CHAINS
cat gpurun_out/r02_gputests_1gpu.log; tail -2 gpurun_out/r02_smoke.log; cat gpurun_out/r02_traffic.log
cut -c1-160 gpurun_out/r02_bench_1gpu.json gpurun_out/r02_bench_dp_1gpu.json
