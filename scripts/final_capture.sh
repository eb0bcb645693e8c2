#!/bin/bash
# Round-end evidence on ONE B200 (run through gpurun from the repo root): the bench line, the reference arm, the ncu
# launch list of the bench command, one `ncu --set full` pass over every kernel worth a capture, the epoch timelines.
# Each ncu pass runs only after the same command has exited 0 without ncu.
set -u
O=gpurun_out
python bench.py > $O/r02_final_bench.json 2> $O/r02_final_bench.err || exit 1
python bench.py --impl reference --steps 5 --warmup 1 > $O/r02_final_ref.json 2> $O/r02_final_ref.err
python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-also --no-roofline > /dev/null 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r02_launches_bench_epoch.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-variants --no-also --no-roofline > $O/r02_ncu_launch.log 2>&1
python scripts/prof_targets.py > /dev/null 2>&1 || exit 3
ncu --set full --clock-control none --import-source on \
    -k regex:"proj_kl|kl_bwd_prep|kl_chol|uniform_|seglik|epoch_|maha|traj_|rsample|head_|gae_|segadv|normalize|adam_|sumsq|tri_inverse" \
    -c 150 -f -o /tmp/r02_full python scripts/prof_targets.py > $O/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > $O/r02_ncu_full_raw.csv 2>> $O/r02_ncu_full.log
ls -la /tmp/r02_full.ncu-rep >> $O/r02_ncu_full.log
python scripts/timeline_epoch.py $O/r02_timeline_epoch_fast.txt > /dev/null 2>&1
tail -c 600 $O/r02_final_bench.json
