#!/bin/bash
# parity subset + update timing + stage trace after a kernel change
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1 || { tail -20 gpurun_out/build.log; exit 1; }
timeout 900 python -X faulthandler -m pytest ${TESTS:-tests} -m gpu -q --maxfail=30 -p no:cacheprovider --timeout=300 > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"
grep -v "^\.\|^$\|^s" gpurun_out/pytest.log | tail -${TAILN:-25} | cut -c1-300
python tools/ab_update.py 300 staged | grep AB_UPDATE
python tools/ab_update.py 300 persistent | grep AB_UPDATE
python tools/shapes_bench.py 2>&1 | tail -4
SACB_TIMELINE=1 timeout 300 python tools/trace_stages.py > gpurun_out/early_trace.txt 2>&1
grep -A5 "stage  [126] \|stage 1[017] " gpurun_out/early_trace.txt | cut -c1-260
grep timeline gpurun_out/early_trace.txt | tail -3
