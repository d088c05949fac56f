set -e
for m in "EPIT 4" "MyEfficientLFNet 4" "DistgSSR 4"; do
  set -- $m
  LFSR_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/ll_$1.csv python profiles/run_minibatch2.py $1 $2 64 > /tmp/ll_$1.log 2>&1 || tail -5 /tmp/ll_$1.log
  python profiles/launch_list.py /tmp/ll_$1.csv > gpurun_out/r02n_$1_launch_summary.csv
  head -30 gpurun_out/r02n_$1_launch_summary.csv
done
