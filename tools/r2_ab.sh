#!/bin/bash
# A/B of experiment builds of the library (tools/bin/libsacb200_<tag>.so) against the default build: update time + per-stage times
set -u
mkdir -p gpurun_out
for v in "" $*; do
  if [ -n "$v" ]; then export SACB_LIB=$PWD/tools/bin/libsacb200_$v.so; else unset SACB_LIB; fi
  echo "== variant ${v:-default}"
  SACB_AB_TAG=${v:-default} python tools/ab_update.py 300 staged | grep AB_UPDATE
  SACB_NOTRACE=1 timeout 300 python tools/trace_stages.py 2>&1 | tail -1
  if [ "${TRACE:-0}" = "1" ]; then
    SACB_TIMELINE=1 timeout 300 python tools/trace_stages.py > gpurun_out/ab_trace_${v:-default}.txt 2>&1
    grep -A4 "stage  1 \|stage 10 " gpurun_out/ab_trace_${v:-default}.txt | cut -c1-250
  fi
done
