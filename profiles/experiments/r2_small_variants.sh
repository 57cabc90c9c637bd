# sim_small_kernel converter variants (C2, fp32 + 16-bit tokens): rows per warp / teams / fp16 stages
for V in base "$@"; do
  if [ $V = base ]; then unset RZ_B200_LIB; else export RZ_B200_LIB=$PWD/radzero_b200/_build/variants/librz_$V.so; fi
  echo "== $V"; python profiles/experiments/r2_time_small.py 2>&1 | tail -4 | head -3
done
