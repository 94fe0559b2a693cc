#!/bin/bash
for v in "" pf8 pf16 pf8mb3 pf16mb3; do
  if [ -z "$v" ]; then lib=""; tag=default; else lib=$PWD/anovos_b200/build/variants/libanovos_b200_$v.so; tag=$v; fi
  if [ -n "$v" ]; then ANOVOS_B200_LIB=$lib python -m pytest tests/test_gpu_kernels.py -x -q -k "moments or histogram" 2>&1 | tail -1; fi
  for nm in mixed; do
    echo "== $tag $nm"
    ANOVOS_B200_LIB=$lib python scripts/kbench.py 1e7 50 $nm 2>&1 | grep -E "^(moments|hist|fused)"
  done
done
