# memcheck of the parity tests that cover every kernel family (small sizes); one sanitizer tool per gpurun call
python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_size and not moments and not philox and not extended" > gpurun_out/san_plain.log 2>&1 || { tail -5 gpurun_out/san_plain.log; exit 1; }
compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -q -m gpu -x -k "not full_size and not moments and not philox and not extended" > gpurun_out/san_memcheck.log 2>&1
echo "memcheck exit $?"; tail -6 gpurun_out/san_memcheck.log
