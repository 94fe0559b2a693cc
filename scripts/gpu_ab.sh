set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/ab_smi.txt 2>&1
timeout 400 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_api.py tests/test_gpu_partitioned.py -x -q > gpurun_out/ab2_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/ab2_pytest.log
: > gpurun_out/ab2_lines.jsonl
timeout 150 python scripts/fused_ab.py 1e8 40 mixed fold_popc >> gpurun_out/ab2_lines.jsonl 2>gpurun_out/ab2_err.log
timeout 150 python scripts/fused_ab.py 1e8 40 none fold_popc >> gpurun_out/ab2_lines.jsonl 2>>gpurun_out/ab2_err.log
timeout 200 python scripts/fused_ab.py 1e8 150 mixed fold_popc_c3 >> gpurun_out/ab2_lines.jsonl 2>>gpurun_out/ab2_err.log
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:scan_kernel -o gpurun_out/r2c_fused python scripts/prof_fused.py 4e7 50 > gpurun_out/ab2_ncu.log 2>&1
cat gpurun_out/ab2_lines.jsonl
tail -3 gpurun_out/ab2_pytest.log
