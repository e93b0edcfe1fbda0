# round-2 ncu evidence: launch list of the bench step, one full capture of the two dominant kernels (and the local GEBV kernel)
set -x
CMD="python bench.py --steps 4 --warmup 3 --no-extras --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"splice_flat_kernel|gebv_counts_kernel" -s 8 -c 2 -o gpurun_out/r2_prof -f $CMD > gpurun_out/ncu2.log 2>&1
python profiles/scripts/local_probe.py 20000 600000 10 0 > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:"local_tc5_kernel" -c 1 -o gpurun_out/r2_prof_local -f python profiles/scripts/local_probe.py 20000 600000 10 0 > gpurun_out/ncu3.log 2>&1
tail -3 gpurun_out/ncu1.log gpurun_out/ncu2.log gpurun_out/ncu3.log
ls -la gpurun_out/
