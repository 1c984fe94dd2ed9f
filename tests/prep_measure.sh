ncu --metrics gpu__time_duration.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:preprocess -s 2 -c 1 --csv --log-file gpurun_out/ncu_prep_tmp.csv python tests/profile_step.py 32 > /dev/null 2>&1; python -c "
import csv
for r in csv.reader(open('gpurun_out/ncu_prep_tmp.csv')):
    if len(r)>10 and 'preprocess' in r[4]: print(r[-3], r[-2], r[-1])"
