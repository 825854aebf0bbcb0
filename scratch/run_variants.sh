for v in s2m3 s4m2 s4m1 s1m4 s1m6; do
  echo "== $v"; BFL_LIB=$PWD/scratch/variants/libbfl_$v.so python tools/time_window.py 2>&1 | tail -1
done
