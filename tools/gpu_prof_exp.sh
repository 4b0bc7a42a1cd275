python tools/exp_solve_mv.py > gpurun_out/exp_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:solve_mv_kernel -s 14 -c 1 -o gpurun_out/prof_c1_noout python tools/exp_solve_mv.py > gpurun_out/ncu_exp.log 2>&1
tail -2 gpurun_out/ncu_exp.log
