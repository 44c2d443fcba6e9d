set -x
CMD="tools/profile_epoch.py --folds 1 --precision f16 --D 12032 --batch 8192 --n-train 16384 --epochs 2"
python $CMD > gpurun_out/c_new.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/c_dp_launches_new.csv python $CMD > /dev/null 2>&1
(cd ab_old && python $CMD > ../gpurun_out/c_old.log 2>&1; ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file ../gpurun_out/c_dp_launches_old.csv python $CMD > /dev/null 2>&1)
cat gpurun_out/c_new.log gpurun_out/c_old.log
