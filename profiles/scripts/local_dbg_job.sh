# knock-outs of local_tc5_kernel (GSB200_LOCAL_DEBUG bits: 1 no operand stores, 2 no mma, 4 no epilogue work, 8 no row loads,
# 16 no B digits) on the two shapes; kernel durations from ncu's launch list (the probe's event span includes host work)
for shape in "20000 600000" "20000 50000"; do
  for d in 0 1 2 4 8 16 7 15 23 31; do
    GSB200_LOCAL_DEBUG=$d timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:local_tc5 --csv python profiles/scripts/local_probe.py $shape 10 1 2>&1 | grep -a "local_tc5_kernel" | tail -1 | awk -F'","' -v d=$d -v s="$shape" '{gsub(/"/,"",$NF); print "shape " s " dbg=" d ": " $NF " " $(NF-1)}'
  done
done > gpurun_out/local_dbg.txt 2>&1
cat gpurun_out/local_dbg.txt
