for wm in 24 32 48 64 96; do for tw in 2 4 8; do echo -n "wide_min=$wm tpc_wide=$tw: "; GPM_WIDE_MIN=$wm GPM_TPC_WIDE=$tw python tools/profile_potrf.py --reps 3 | tail -1; done; done
for tn in 4 16; do echo -n "tpc_narrow=$tn: "; GPM_TPC_NARROW=$tn python tools/profile_potrf.py --reps 3 | tail -1; done
echo "N=8192:"; for wm in 16 32 48 1000; do echo -n "wide_min=$wm: "; GPM_WIDE_MIN=$wm python tools/profile_potrf.py --n 8192 --reps 3 | tail -1; done
