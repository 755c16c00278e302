mkdir -p gpurun_out
timeout 900 python -X faulthandler -m pytest tests/test_gpu_replay.py -m gpu -q -x -p no:cacheprovider --timeout=300 > gpurun_out/quick_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 gpurun_out/quick_pytest.log
python tools/per_profile.py 1000000 300 0
python tools/per_profile.py 100000 300 0
ncu --metrics gpu__time_duration.sum --clock-control none --cache-control none -k regex:per_ -s 30 -c 6 --csv --log-file gpurun_out/per_launches.csv python tools/per_profile.py 1000000 10 0 > gpurun_out/per_ncu.log 2>&1
python - <<P
import csv
rows=[r for r in csv.reader(open("gpurun_out/per_launches.csv")) if len(r)>5]
hdr=rows[0]; ki=hdr.index("Kernel Name"); vi=hdr.index("Metric Value"); gi=hdr.index("Grid Size") if "Grid Size" in hdr else None
for r in rows[1:]: print(r[ki][:40], r[vi], r[gi] if gi is not None else "")
P
ncu --set full --clock-control none --cache-control none --import-source on -k regex:per_ -s 30 -c 3 -f -o gpurun_out/per_full python tools/per_profile.py 1000000 10 0 > gpurun_out/per_ncu2.log 2>&1
echo "full rc=$?"
