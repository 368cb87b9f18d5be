#!/bin/bash
# dev helper: run the B200-bound bwa mem on synthetic data with the job dump on, then replay the dump
# usage: harvest_jobs.sh <pe|se> <reads> <len> <sub> <indel> <indel_max>
D=$(mktemp -d); cd /root/repo
python - <<PY
import sys; sys.path.insert(0,'tests')
import samtest as S
g=S.write_genome("$D/ref.fa", 5000000, seed=1)
S.bwa_index("$D/ref.fa")
paths=["$D/r1.fq","$D/r2.fq"] if "$1"=="pe" else ["$D/r1.fq"]
S.write_reads_fast(paths, g, $2, $3, seed=2, sub=$4, indel=$5, indel_max=$6)
PY
READS="$D/r1.fq"; [ "$1" = pe ] && READS="$D/r1.fq $D/r2.fq"
KSW_B200_REF=0 KSW_B200_DUMP=$D/jobs integration/_bin/bwa_b200 mem -t 1 -b 1000000 $D/ref.fa $READS > /dev/null 2> $D/err; tail -2 $D/err
ls -la $D/*.bin
python scripts/bench_jobs.py $D/jobs.*.bin
rm -rf $D
