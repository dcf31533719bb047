for cfg in "48 64 96" "32 64 96" "24 48 96" "32 48 80" "24 40 64" "16 32 64" "40 56 80" "64 96 128"; do
  set -- $cfg
  for n in 16384 8192; do
    r=$(GPM_WIDE_MIN=$1 GPM_WIDE4_MIN=$2 GPM_WIDE8_MIN=$3 python tools/profile_potrf.py --n $n --reps 4 | tail -2 | awk '{print $5}' | tr '\n' ' ')
    echo "wide=$1 wide4=$2 wide8=$3 N=$n: $r"
  done
done
