# N = 3072 / 4096: outer-panel width threshold and tiles per CTA of the narrow look-ahead updates, re-swept after the
# diagonal-block kernel got 10 % shorter and the look-ahead column was split (round 2)
for n in 3072 4096; do
  for cfg in "36 8" "32 8" "28 8" "24 8" "36 4" "36 2" "36 16" "32 4"; do
    set -- $cfg
    r=$(GPM_WIDE_MIN=$1 GPM_TPC_NARROW=$2 python tools/profile_potrf.py --n $n --reps 6 | tail -3 | awk '{print $5}' | tr '\n' ' ')
    echo "N=$n wide_min=$1 tpc_narrow=$2: $r"
  done
done
