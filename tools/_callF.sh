set -x
CMD="tools/profile_epoch.py --folds 1 --precision f16 --D 12032 --batch 8192 --n-train 16384 --epochs 2"
python $CMD > gpurun_out/f_new.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/f_dp_launches.csv python $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_prep -c 1 -o /tmp/f_prep python $CMD > /dev/null 2>&1
python profiles/extract_ncu.py /tmp/f_prep.ncu-rep 0 > gpurun_out/f_prep.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_gemm_tc -s 4 -c 1 -o /tmp/f_g3 python $CMD > /dev/null 2>&1
python profiles/extract_ncu.py /tmp/f_g3.ncu-rep 0 > gpurun_out/f_g3.txt 2>&1
cat gpurun_out/f_new.log
