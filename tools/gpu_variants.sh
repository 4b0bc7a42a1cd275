for v in "" build/variants/bl15000.so; do
  echo "== ${v:-default}"; RODEO_B200_LIB=$v python tools/exp_solve_mv.py > /tmp/o.txt 2>&1; cat /tmp/o.txt | tail -3
done
