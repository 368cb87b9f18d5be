#!/bin/bash
# Development helper (GPU box): times the resident kernel path for the in-tree library and every A/B variant built by
# scripts/build_variant.sh.   scripts/ab_variants.sh [jobs]
n=${1:-2000000}
cd "$(dirname "$0")/.."
echo "== default"; timeout 300 python scripts/prof_fast.py $n 5 2>&1 | tail -1
for f in bwa_mem_quickassist_b200/build/variants/libksw_b200_*.so; do
  echo "== $f"; KSW_B200_LIB=$PWD/$f timeout 300 python scripts/prof_fast.py $n 5 2>&1 | tail -1
done
