for v in "" build/variants/st3_4000.so build/variants/st6_2000.so build/variants/st4_8000.so; do
  echo "== ${v:-default}"; RODEO_B200_LIB=$v python tools/exp_solve_mv.py > /tmp/o.txt 2>&1; head -1 /tmp/o.txt
done
