mkdir -p gpurun_out
for v in 0 1; do
  if [ $v = 1 ]; then export SACB_FIXED_TILES=1; fi
  python bench.py --steps 300 --warmup 20 --no-cpu > gpurun_out/ab_$v.log 2>&1
  python - <<P
import json
for l in open("gpurun_out/ab_$v.log"):
    if l.startswith("{"):
        d=json.loads(l); print("fixed=$v", d["ms_per_step"], d["roofline"]["ms_per_launch_sum"], d["e2e"]["value"], d["roofline"]["per_sample"]["ms_per_call"])
P
done
