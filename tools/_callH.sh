set -x
for ch in 2 4 8; do
MRGAN_CHAINS=$ch python bench.py --folds 37 --steps 4 --warmup 3 --no-cpu --no-modes > gpurun_out/h_f37_ch$ch.json 2> gpurun_out/h_f37_ch$ch.err
done
for ch in 4 8; do
MRGAN_CHAINS=$ch python tools/table1_rank_probe.py 4 8 > gpurun_out/h_probe_ch$ch.log 2>&1
done
cut -c1-140 gpurun_out/h_f37_ch*.json
grep -h "per epoch" gpurun_out/h_probe_ch*.log
