#!/bin/bash
# Round-end evidence on ONE B200 (run through gpurun from the repo root, AFTER scripts/capture_ncu_full.sh has been
# summarised into profiles/r02_ncu_traffic.json so that the bench line carries the measured traffic): the bench line,
# the reference arm, the ncu launch list of the bench command (only after the same command exited 0 without ncu),
# the epoch timeline and the HBM-family tables.
set -u
O=gpurun_out
python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_final_ref.json 2> $O/r02_final_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-also --no-roofline > /dev/null 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench_epoch.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-also --no-roofline > $O/r02_ncu_launch.log 2>&1
python scripts/timeline_epoch.py $O/r02_timeline_epoch_fast.txt > /dev/null 2>&1
for b in 65536 16384 1024; do python scripts/hbm_kernels.py $b $O/r02_final_hbm_$b.txt > /dev/null 2>&1; done
tail -c 600 $O/r02_final_bench.json
