# one B200: per-kernel ncu metrics of ONE eager conv step per workload (bench.py --profile-step brackets it
# with cudaProfilerStart/Stop), reduced on the box to the text / json kept under profiles/ (the .ncu-rep
# files are too large to bring back)
mkdir -p gpurun_out
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,sm__throughput.avg.pct_of_peak_sustained_elapsed,l1tex__throughput.avg.pct_of_peak_sustained_elapsed,lts__throughput.avg.pct_of_peak_sustained_elapsed
for w in ${WORKLOADS:-cikm amazon-full}; do
  ncu --metrics $M --clock-control none --profile-from-start off -o /tmp/ncu_$w -f python bench.py --workload $w --also none --steps 3 --warmup 3 --no-cpu-baseline --no-graph --profile-step > gpurun_out/${TAG:-r2}_ncu_$w.log 2>&1
  python profiles/ncu_extract.py /tmp/ncu_$w.ncu-rep gpurun_out/${TAG:-r2}_ncu_$w.txt --traffic gpurun_out/${TAG:-r2}_ncu_traffic_$w.json --workload $w --steps-captured 1 --how "--metrics <list in profiles/scripts/r02_run_ncu.sh>" --note "one eager conv step of $w (bench.py --profile-step)" > /dev/null 2>> gpurun_out/${TAG:-r2}_ncu_$w.log
  ncu -i /tmp/ncu_$w.ncu-rep --page raw --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; c={k:i for i,k in enumerate(h)}
print('id,kernel,duration(' + rows[1][c['gpu__time_duration.sum']] + ')')
for r in rows[2:]:
    print(r[c['ID']] + ',' + r[c['Kernel Name']].split('(')[0].replace('void ','') + ',' + r[c['gpu__time_duration.sum']])
" > gpurun_out/${TAG:-r2}_launches_$w.csv
done
ls -la gpurun_out /tmp/*.ncu-rep
