mkdir -p gpurun_out
SACB_TIMELINE=1 python tools/trace_stages.py > gpurun_out/timeline.log 2>&1
grep timeline gpurun_out/timeline.log
