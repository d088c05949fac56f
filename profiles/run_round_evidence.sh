# Round evidence in one GPU call (one B200): GPU tests, the bench line, per-network launch lists, the kernel zoo
# (plain timings + one `ncu --set full` capture per kernel). Everything lands in gpurun_out/ with the tag given as $1.
TAG=${1:-r02s}
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_gputests.log 2>&1; tail -2 gpurun_out/${TAG}_gputests.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 300 gpurun_out/${TAG}_bench.json; echo
for m in "EPIT 4" "MyEfficientLFNet 4" "DistgSSR 4"; do
  set -- $m
  LFSR_CUDA_GRAPH=0 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file /tmp/ll_$1.csv python profiles/run_minibatch2.py $1 $2 64 > /tmp/ll_$1.log 2>&1 || tail -5 /tmp/ll_$1.log
  python profiles/launch_list.py /tmp/ll_$1.csv > gpurun_out/${TAG}_$1_launch_summary.csv
  grep -h "patches/s" /tmp/ll_$1.log | tail -1
done
python profiles/run_kernel_zoo.py 64 > gpurun_out/${TAG}_kernel_zoo_plain.log 2>&1; tail -30 gpurun_out/${TAG}_kernel_zoo_plain.log
LFSR_ZOO_REPS=1 ncu --set full --clock-control none -k "regex:^(block_mean|conv_|divide|dw_tile|sa_tile|ang_expand|pooled_mlp|integrate|interp|layernorm|mel_epi|metric)" -c 60 -o /tmp/zoo python profiles/run_kernel_zoo.py 64 > /tmp/zoo.log 2>&1 || tail -5 /tmp/zoo.log
ncu -i /tmp/zoo.ncu-rep --page raw --csv > gpurun_out/${TAG}_zoo_raw.csv 2>/dev/null
ls -la gpurun_out/${TAG}_*
