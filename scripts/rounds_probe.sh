#!/bin/bash
D=$(mktemp -d); cd /root/repo
python - <<PY
import sys; sys.path.insert(0,'tests')
import samtest as S
g=S.write_genome("$D/ref.fa", 1000000, seed=1)
S.bwa_index("$D/ref.fa")
S.write_reads_fast(["$D/r.fq"], g, ${1:-1000000}, 100, seed=2)
PY
KSW_B200_SCHED=rounds integration/_bin/bwa_b200 mem -t ${2:-16} $D/ref.fa $D/r.fq > $D/b.sam 2> $D/b.err; echo "rc=$?"; grep -v "^\[M::main_mem\] read" $D/b.err | head -12
rm -rf $D
