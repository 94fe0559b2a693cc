set -x
mkdir -p gpurun_out
timeout 780 python bench.py > gpurun_out/r2b_bench_c3.json 2> gpurun_out/r2b_bench_c3.err; echo "bench rc=$?"
tail -c 600 gpurun_out/r2b_bench_c3.err
: > gpurun_out/ab3_lines.jsonl
timeout 200 python scripts/fused_ab.py 1e8 150 mixed hybrid_c3 >> gpurun_out/ab3_lines.jsonl 2>gpurun_out/ab3_err.log
timeout 100 python scripts/fused_ab.py 1e8 40 none hybrid >> gpurun_out/ab3_lines.jsonl 2>>gpurun_out/ab3_err.log
cat gpurun_out/ab3_lines.jsonl
