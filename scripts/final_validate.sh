# round-end evidence on one B200: bench line, reference arm, ncu launch list of a step and per-kernel counters
# (-> profiles/ via gpurun_out/); the .ncu-rep itself (70 MB) stays on the box
set -x
timeout -s KILL 300 python bench.py 2>/dev/null | grep "^{" > gpurun_out/r2_final_n1.json; cut -c1-300 gpurun_out/r2_final_n1.json
timeout -s KILL 400 python bench.py --impl reference 2>/dev/null | grep "^{" > gpurun_out/r2_final_ref.json; cut -c1-200 gpurun_out/r2_final_ref.json
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_final_launches.csv python bench.py --steps 2 --warmup 1 --graph 0 > /dev/null 2>&1
timeout -s KILL 600 ncu --set full --clock-control none -k regex:"k_" -s 60 -c 34 -f -o /tmp/r2_final_step python bench.py --steps 2 --warmup 1 --graph 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log
python scripts/ncu_counters.py /tmp/r2_final_step.ncu-rep gpurun_out/r2_counters.json "python bench.py --steps 2 --warmup 1 --graph 0" | tee gpurun_out/r2_counters.txt
ncu -i /tmp/r2_final_step.ncu-rep --page details --csv 2>/dev/null | python -c "
import csv,sys
rows=list(csv.reader(sys.stdin)); h=rows[0]; kn=h.index('Kernel Name'); mn=h.index('Metric Name'); mv=h.index('Metric Value')
keep=('Duration','Registers Per Thread','Achieved Occupancy','Theoretical Occupancy','Executed Ipc Active','Issue Slots Busy','DRAM Throughput','L2 Cache Throughput','L1/TEX Cache Throughput','Compute (SM) Throughput','Memory Throughput','Dynamic Shared Memory Per Block','Block Limit Registers','Block Limit Shared Mem','Grid Size','Block Size')
for r in rows[1:]:
    if r[mn] in keep: print(r[0], r[kn].split('(')[0][:44], '|', r[mn], '|', r[mv])
" > gpurun_out/r2_final_details.txt
timeout -s KILL 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_final_launches_rank_of_8.csv python scripts/slab_profile.py 8 > /dev/null 2>&1
for cfg in c3 c5 c2; do timeout -s KILL 300 python bench.py --config $cfg 2>/dev/null | grep "^{" > gpurun_out/r2_final_$cfg.json; cut -c1-200 gpurun_out/r2_final_$cfg.json; done
du -sh gpurun_out
