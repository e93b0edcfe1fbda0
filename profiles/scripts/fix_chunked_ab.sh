for f in 0 1; do echo "GSB200_SPLICE_FIX=$f"; GSB200_SPLICE_FIX=$f python profiles/scripts/config4_probe.py 500000 100000 both; done
echo "headline, forced on/off"
for f in 0 1; do GSB200_SPLICE_FIX=$f python bench.py --no-extras --no-cpu-baseline --steps 60 > gpurun_out/fix$f.json 2>/dev/null; python -c "
import json
d=json.loads(open('gpurun_out/fix$f.json').read().strip().splitlines()[-1])
print('fix $f', d['ms_per_step'], d['phases_ms']['pairs_and_splice_ms'])"; done
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
GSB200_SPLICE_FIX=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -2
