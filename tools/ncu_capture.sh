#!/bin/bash
# ncu evidence of one bench workload (run under gpurun, one workload per call -- all ncu runs of a call count as one):
#   tools/ncu_capture.sh <tag> <kernel regex> <launch skip> <bench.py arguments...>
# 1. the plain command (must exit 0), 2. launch list with device times, 3. --set full capture of ONE launch of the kernel,
# exported as raw CSV.  Outputs: gpurun_out/<tag>_launches.csv, gpurun_out/<tag>_full.raw.csv (+ .ncu-rep)
tag=$1; kre=$2; skip=$3; shift 3
cmd="python bench.py $* --no-also --no-cpu-baseline --no-e2e --no-parity-gate"
$cmd > gpurun_out/${tag}_plain.log 2>&1 || { echo "plain run failed"; tail -5 gpurun_out/${tag}_plain.log; exit 1; }
tail -c 400 gpurun_out/${tag}_plain.log; echo
ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/${tag}_launches.csv $cmd > gpurun_out/${tag}_ncu1.log 2>&1
echo "launch list rc=$? lines=$(wc -l < gpurun_out/${tag}_launches.csv)"
ncu --set full --clock-control none --import-source on -k regex:$kre -s $skip -c 1 -f -o gpurun_out/${tag}_full $cmd > gpurun_out/${tag}_ncu2.log 2>&1
echo "full capture rc=$?"; tail -3 gpurun_out/${tag}_ncu2.log
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv > gpurun_out/${tag}_full.raw.csv 2>/dev/null
ncu -i gpurun_out/${tag}_full.ncu-rep --page raw --csv 2>/dev/null | python3 -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; v=rows[-1]
for k in ('Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_subpipe_dmma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_shared_cycles_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','launch__grid_size','sm__warps_active.avg.per_cycle_active','lts__t_bytes.sum','dram__throughput.avg.pct_of_peak_sustained_elapsed'):
    if k in h: print(k, rows[1][h.index(k)], v[h.index(k)])
"
ls -la gpurun_out/${tag}_full.ncu-rep
