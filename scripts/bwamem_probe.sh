#!/bin/bash
# dev helper: time stock vs b200 bwa mem on a larger SE set with traces
D=$(mktemp -d); cd /root/repo
python - <<PY
import sys; sys.path.insert(0,'tests')
import samtest as S
g=S.write_genome("$D/ref.fa", 2000000, seed=1)
S.bwa_index("$D/ref.fa")
S.write_reads_se("$D/r.fq", g, ${1:-600000}, 100, seed=2)
PY
T=${2:-16}
S=$(date +%s.%N); oracle/_ref/bwa_stock mem -t $T $D/ref.fa $D/r.fq > $D/s.sam 2> $D/s.err; E=$(date +%s.%N); echo "stock wall $(python -c "print(round($E - $S, 3))") s"; tail -2 $D/s.err
for B in 1 100000; do
S=$(date +%s.%N); integration/_bin/bwa_b200 mem -t $T -b $B $D/ref.fa $D/r.fq > $D/b.sam 2> $D/b.err || { tail -3 $D/b.err; }; E=$(date +%s.%N); echo "b200 -b $B wall $(python -c "print(round($E - $S, 3))") s"; grep -E "Processed|Real" $D/b.err | tail -4
cmp <(grep -v '^@PG' $D/s.sam) <(grep -v '^@PG' $D/b.sam) && echo identical
done
rm -rf $D
