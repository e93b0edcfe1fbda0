# knock-outs of local_tc5_kernel (GSB200_LOCAL_DEBUG bits: 1 no operand stores, 2 no mma, 4 no epilogue work, 8 no row loads,
# 16 no B digits, 64 no output stores) on the two shapes: device span of the probe
for shape in "20000 600000" "20000 50000"; do
  for d in 0 1 2 4 8 16 64 68; do GSB200_LOCAL_DEBUG=$d timeout 300 python profiles/scripts/local_probe.py $shape 10 2 2>&1 | grep -a "path=\|rror" | sed "s/^/dbg=$d /"; done
done > gpurun_out/local_dbg.txt 2>&1
cat gpurun_out/local_dbg.txt
