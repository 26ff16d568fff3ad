#!/bin/bash
# ncu evidence of round 2 (run under gpurun on one B200): launch list of ONE 2^20 proof, and --set full captures of the
# dominant kernels (plain and TMA-staged NTT pass side by side).  Raw pages are exported to CSV on the box.
set -x
OUT=gpurun_out
REP=/tmp/ncu_r02   # the .ncu-rep files stay on the box (gpurun merges at most 64 MiB back): only CSV exports travel
mkdir -p $REP
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file $OUT/launches_r02.csv python tools/profile_prove.py 20 > $OUT/ncu_launches_r02.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'msm_accum_kernel|msm_fine_sort_kernel|msm_coarse_kernel|quotient_kernel' -c 10 -o $REP/prof_msm_r02 -f python tools/profile_prove.py 20 > $OUT/ncu_msm_r02.log 2>&1
ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ntt_pass_kernel' -s 4 -c 6 -o $REP/prof_ntt_r02 -f python tools/profile_prove.py 20 > $OUT/ncu_ntt_r02.log 2>&1
PK_NTT_TMA=1 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'ntt_pass_tma_kernel' -s 4 -c 6 -o $REP/prof_ntt_tma_r02 -f python tools/profile_prove.py 20 > $OUT/ncu_ntt_tma_r02.log 2>&1
for f in prof_msm_r02 prof_ntt_r02 prof_ntt_tma_r02; do ncu -i $REP/$f.ncu-rep --page raw --csv > $OUT/raw_$f.csv 2>/dev/null; done
ls -la $REP/*.ncu-rep $OUT/raw_prof_*r02.csv
