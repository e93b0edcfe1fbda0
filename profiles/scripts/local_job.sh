set -x
for p in mma tc5; do GSB200_LOCAL_PATH=$p timeout 300 python profiles/scripts/local_probe.py 20000 50000 1,4,5,8,10 3; done > gpurun_out/local_probe.txt 2>&1
for p in mma tc5; do GSB200_LOCAL_PATH=$p timeout 300 python profiles/scripts/local_probe.py 20000 600000 10 2; done >> gpurun_out/local_probe.txt 2>&1
cat gpurun_out/local_probe.txt | grep path=
GSB200_LOCAL_PATH=tc5 timeout 600 ncu --set full --import-source on --clock-control none -k regex:local_tc5 -c 1 -o gpurun_out/local_tc5_probe python profiles/scripts/local_probe.py 20000 50000 10 0 > gpurun_out/ncu_local.log 2>&1
tail -3 gpurun_out/ncu_local.log
ls -la gpurun_out/*.ncu-rep
