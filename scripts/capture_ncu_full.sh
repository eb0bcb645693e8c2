#!/bin/bash
# `ncu --set full` over every kernel worth a capture (scripts/prof_targets.py: two headline epochs, two epochs with
# per-episode covariances, the HBM family at B = 16384), after the same command has exited 0 without ncu.
# -> gpurun_out/r02_ncu_full_raw.csv (summarised here by scripts/ncu_summary.py into profiles/)
set -u
O=gpurun_out
python scripts/prof_targets.py > $O/r02_prof_targets.log 2>&1 || exit 3
ncu --set full --clock-control none --import-source on \
    -k regex:"proj_kl|kl_bwd_prep|kl_chol|uniform_|seglik|epoch_|maha|traj_|rsample|head_|gae_|segadv|normalize|adam_|sumsq|tri_inverse" \
    -c 110 -f -o /tmp/r02_full python scripts/prof_targets.py > $O/r02_ncu_full.log 2>&1
ncu -i /tmp/r02_full.ncu-rep --page raw --csv > $O/r02_ncu_full_raw.csv 2>> $O/r02_ncu_full.log
ls -la /tmp/r02_full.ncu-rep >> $O/r02_ncu_full.log
tail -3 $O/r02_ncu_full.log
