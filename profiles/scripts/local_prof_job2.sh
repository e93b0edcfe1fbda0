rm -f gpurun_out/local_prof.txt
for shape in "20000 600000" "20000 50000"; do
for d in 0 2; do
    echo "shape $shape dbg=$d" >> gpurun_out/local_prof.txt
    GSB200_LOCAL_PROF=gpurun_out/local_prof.txt GSB200_LOCAL_DEBUG=$d timeout 300 python profiles/scripts/local_probe.py $shape 10 0 2>&1 | grep -a "rror"
done
done
cat gpurun_out/local_prof.txt
