set -x
CMD="tools/profile_epoch.py --folds 1 --precision f16 --D 12032 --batch 8192 --n-train 16384 --epochs 1"
ncu --set full --clock-control none --import-source on -k regex:k_prep -c 1 -o /tmp/d_prep_new python $CMD > gpurun_out/d_new.log 2>&1
(cd ab_old && ncu --set full --clock-control none --import-source on -k regex:k_prep -c 1 -o /tmp/d_prep_old python $CMD > ../gpurun_out/d_old.log 2>&1)
ls -la /tmp/*.ncu-rep
python profiles/extract_ncu.py /tmp/d_prep_new.ncu-rep 0 > gpurun_out/d_prep_new.txt 2>&1
python profiles/extract_ncu.py /tmp/d_prep_old.ncu-rep 0 > gpurun_out/d_prep_old.txt 2>&1
ncu -i /tmp/d_prep_old.ncu-rep --page details > gpurun_out/d_prep_old_details.txt 2>&1
ncu -i /tmp/d_prep_new.ncu-rep --page details > gpurun_out/d_prep_new_details.txt 2>&1
for f in /tmp/d_prep_*.ncu-rep; do s=$(stat -c %s $f); if [ $s -lt 25000000 ]; then cp $f gpurun_out/; fi; done
