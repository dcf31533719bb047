# N = 2048 / 4096 / 6144: outer-panel width thresholds and the no-look-ahead variant (re-sweep after the strip kernel)
for n in 2048 4096 6144; do
  for cfg in "32 64 96" "24 64 96" "16 64 96" "12 64 96" "8 64 96" "16 24 96" "1000 1000 1000"; do
    set -- $cfg
    r=$(GPM_WIDE_MIN=$1 GPM_WIDE4_MIN=$2 GPM_WIDE8_MIN=$3 python tools/profile_potrf.py --n $n --reps 6 | tail -3 | awk '{print $5}' | tr '\n' ' ')
    echo "N=$n wide=$1 wide4=$2 wide8=$3: $r"
  done
  r=$(GPM_NO_LOOKAHEAD=1 python tools/profile_potrf.py --n $n --reps 6 | tail -3 | awk '{print $5}' | tr '\n' ' ')
  echo "N=$n no_lookahead: $r"
done
