#!/bin/bash
# Run on the GPU box (via gpurun): plain run first, then the ncu launch list and one full capture
# of the three EVM kernels of the same command.  Outputs land in gpurun_out/ with prefix $1.
set -u
P=${1:-prof}
CMD="python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu --resident 1"
mkdir -p gpurun_out
$CMD > gpurun_out/${P}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -s 19 -c 12 --csv --log-file gpurun_out/${P}_launches.csv $CMD > gpurun_out/${P}_ncu_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:pyrdown|bandpass|collapse' -s 9 -c 3 -f -o gpurun_out/${P}_full $CMD > gpurun_out/${P}_ncu_full.log 2>&1
tail -2 gpurun_out/${P}_plain.log | cut -c1-600
tail -3 gpurun_out/${P}_ncu_full.log
