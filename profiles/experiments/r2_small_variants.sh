# sim_small_kernel with two rows per converter warp: team / stage / tile-width variants (C2, fp32 + 16-bit tokens)
for V in base t4 t8s4 t8s3 tok32; do
  if [ $V = base ]; then unset RZ_B200_LIB; else export RZ_B200_LIB=$PWD/radzero_b200/_build/variants/librz_$V.so; fi
  echo "== $V"; python profiles/experiments/r2_time_small.py 2>&1 | tail -4 | head -3
done
