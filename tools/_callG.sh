set -x
python tools/table1_rank_probe.py 3 8 > gpurun_out/g_probe8.log 2>&1
python tools/table1_rank_probe.py 3 4 > gpurun_out/g_probe4.log 2>&1
tail -5 gpurun_out/g_probe8.log gpurun_out/g_probe4.log
