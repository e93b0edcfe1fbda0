rm -f gpurun_out/local_prof.txt
for d in 63 32; do
    echo "shape 20000 600000 dbg=$d (issuer: wait(operand) := step loops, wait(accumulator free) := unit commits)" >> gpurun_out/local_prof.txt
    GSB200_LOCAL_PROF=gpurun_out/local_prof.txt GSB200_LOCAL_DEBUG=$d timeout 300 python profiles/scripts/local_probe.py 20000 600000 10 0 2>&1 | grep -a "rror"
done
cat gpurun_out/local_prof.txt
timeout 300 python -m pytest tests/test_gpu_local_tc5.py -x -q 2>&1 | grep -a "assert\|Error\|error\|^E " | head -20
