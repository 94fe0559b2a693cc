#!/bin/bash
for v in "" u8; do
  if [ -z "$v" ]; then lib=""; tag=default; else lib=$PWD/anovos_b200/build/variants/libanovos_b200_$v.so; tag=$v; fi
  for nm in none mixed; do
    echo "== $tag $nm"
    ANOVOS_B200_LIB=$lib python scripts/kbench.py 1e7 50 $nm 2>&1 | grep -E "^(moments|hist|fused)"
  done
done
