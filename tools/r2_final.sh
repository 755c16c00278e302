#!/bin/bash
# round-2 evidence run (1 GPU): full GPU test suite, smoke, bench line, ncu launch list of the bench command, stage lists
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -X faulthandler -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 300 --warmup 20 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -3 gpurun_out/bench.err
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref.log 2>&1; echo "ref rc=$?"
CMD="python bench.py --steps 12 --warmup 3 --no-cpu --no-sharded"
$CMD > gpurun_out/plain_bench.log 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --cache-control none -c 700 \
    --csv --log-file gpurun_out/r02_bench_launch_list.csv $CMD > gpurun_out/ncu_launches.log 2>&1
echo "launch list rc=$?"
python tools/ab_update.py 300 staged | grep AB_UPDATE
python tools/ab_update.py 300 persistent | grep AB_UPDATE
SACB_TIMELINE=1 timeout 300 python tools/trace_stages.py > gpurun_out/r02_stage_trace.txt 2>&1
python tools/shapes_bench.py 2>&1 | tail -6
python tools/throughput_stages.py 1 8192 2>&1 | grep "THROUGHPUT\|stage_us"
python tools/throughput_stages.py 128 256 2>&1 | grep "THROUGHPUT\|stage_us"
python tools/act_latency.py 2>&1 | tail -2
