mkdir -p gpurun_out
for cap in 0 -1 296; do
  export SACB_GRID_CAP=$cap
  [ $cap = -1 ] && unset SACB_GRID_CAP
  for a in 8 64; do
    echo "cap=$cap"; timeout 300 python tools/population_bench.py $a 10 2>&1 | tail -1
  done
done
unset SACB_GRID_CAP
timeout 600 python -m pytest tests/test_gpu_sharded.py -m gpu -q -x -p no:cacheprovider 2>&1 | tail -3
