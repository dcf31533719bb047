for w4 in 40 48 56 64; do for w2 in 24 32 48; do echo -n "wide4_min=$w4 wide_min=$w2: "; GPM_WIDE4_MIN=$w4 GPM_WIDE_MIN=$w2 python tools/profile_potrf.py --reps 3 | tail -1; done; done
for w8 in 96 112; do echo -n "wide8_min=$w8 (w4=56,w2=32): "; GPM_WIDE8_MIN=$w8 GPM_WIDE4_MIN=56 GPM_WIDE_MIN=32 python tools/profile_potrf.py --reps 3 | tail -1; done
echo "N=8192:"; for w4 in 40 48 64 1000; do echo -n "wide4_min=$w4 wide_min=32: "; GPM_WIDE4_MIN=$w4 GPM_WIDE_MIN=32 python tools/profile_potrf.py --n 8192 --reps 3 | tail -1; done
