set -x
mkdir -p gpurun_out
# (0) the sort tests on the default build (scatter copy-out changed since the last GPU run)
timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "mode or sort or distinct or partition or hll or onesweep" > gpurun_out/fin_sort_tests.log 2>&1; echo "rc=$?" >> gpurun_out/fin_sort_tests.log
tail -2 gpurun_out/fin_sort_tests.log
# (1) pack variants: timing + checksums
: > gpurun_out/fin_sort_ab.jsonl
timeout 120 python scripts/sort_ab.py 1e8 12 default >> gpurun_out/fin_sort_ab.jsonl 2>gpurun_out/fin_sort_ab.err
for v in ps1 ps5 ps6; do
  ANOVOS_B200_LIB=$PWD/anovos_b200/build/variants/libanovos_b200_$v.so timeout 120 python scripts/sort_ab.py 1e8 12 $v >> gpurun_out/fin_sort_ab.jsonl 2>>gpurun_out/fin_sort_ab.err
done
timeout 120 python scripts/sort_ab.py 1e8 12 default_again >> gpurun_out/fin_sort_ab.jsonl 2>>gpurun_out/fin_sort_ab.err
cat gpurun_out/fin_sort_ab.jsonl
ANOVOS_B200_LIB=$PWD/anovos_b200/build/variants/libanovos_b200_ps5.so timeout 200 python -m pytest tests/test_gpu_kernels.py -x -q -k "mode or sort or distinct or partition or hll or onesweep" > gpurun_out/fin_sort_tests_ps5.log 2>&1; echo "rc=$?" >> gpurun_out/fin_sort_tests_ps5.log
tail -2 gpurun_out/fin_sort_tests_ps5.log
# (2) fused pass: burst + sustained with NVML clocks, and the ncu capture of the final kernels
timeout 200 python scripts/fused_ab.py 1e8 150 mixed final_c3 > gpurun_out/fin_fused_ab.jsonl 2>gpurun_out/fin_fused_ab.err
cat gpurun_out/fin_fused_ab.jsonl
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_kernel -o gpurun_out/r2d_fused python scripts/prof_fused.py 4e7 50 > gpurun_out/fin_ncu.log 2>&1
# (3) the rest of the GPU suite, as far as the budget goes
timeout 330 python -m pytest tests -x -q -m gpu > gpurun_out/fin_pytest_gpu.log 2>&1; echo "rc=$?" >> gpurun_out/fin_pytest_gpu.log
tail -3 gpurun_out/fin_pytest_gpu.log
