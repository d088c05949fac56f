set -e
python profiles/run_sa.py 64 > gpurun_out/sa_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:sa_tile_kernel --launch-skip 5 --launch-count 1 -o /tmp/sa python profiles/run_sa.py 64 > /tmp/ncu_sa.log 2>&1 || tail -5 /tmp/ncu_sa.log
ncu -i /tmp/sa.ncu-rep --page source --csv > gpurun_out/sa_source.csv 2>/dev/null
ncu -i /tmp/sa.ncu-rep --page raw --csv > gpurun_out/sa_raw.csv 2>/dev/null
ls -la gpurun_out/sa_*; cat gpurun_out/sa_plain.log
