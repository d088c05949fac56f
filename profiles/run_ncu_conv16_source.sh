set -e
python profiles/run_conv16.py 64 64 3 5 64 160 > gpurun_out/conv16_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:conv_tc_kernel --launch-skip 30 --launch-count 1 -o /tmp/c16 python profiles/run_conv16.py 64 64 3 5 64 160 > /tmp/ncu_c16.log 2>&1 || tail -5 /tmp/ncu_c16.log
ncu -i /tmp/c16.ncu-rep --page source --csv > gpurun_out/c16_source.csv 2>/dev/null
ncu -i /tmp/c16.ncu-rep --page raw --csv > gpurun_out/c16_raw.csv 2>/dev/null
ls -la gpurun_out/c16_*; cat gpurun_out/conv16_plain.log
