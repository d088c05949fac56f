# one --set full capture (with source) of an EPI-branch kernel at the Track-2 shape, batch 64. usage: run_ncu_epi.sh [kernel regex]
set -e
K=${1:-mel_epi_branch_mma_kernel}
python profiles/run_epi.py 64 both > gpurun_out/epi_plain.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:$K --launch-skip 5 --launch-count 1 -o /tmp/epi python profiles/run_epi.py 64 tc > /tmp/ncu_epi.log 2>&1 || tail -5 /tmp/ncu_epi.log
ncu -i /tmp/epi.ncu-rep --page source --csv > gpurun_out/epi_source.csv 2>/dev/null
ncu -i /tmp/epi.ncu-rep --page raw --csv > gpurun_out/epi_raw.csv 2>/dev/null
ls -la gpurun_out/epi_*; cat gpurun_out/epi_plain.log
