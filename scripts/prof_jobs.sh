#!/bin/bash
# dev helper: harvest a real job mix (PE150) and capture an ncu full profile of the fast kernel on it
D=$(mktemp -d); cd /root/repo
python - <<PY
import sys; sys.path.insert(0,'tests')
import samtest as S
g=S.write_genome("$D/ref.fa", 5000000, seed=1)
S.bwa_index("$D/ref.fa")
S.write_reads_fast(["$D/r1.fq","$D/r2.fq"], g, 100000, 150, seed=2, sub=0.01, indel=0.001, indel_max=1)
PY
KSW_B200_REF=0 KSW_B200_DUMP=$D/jobs integration/_bin/bwa_b200 mem -t 1 -b 1000000 $D/ref.fa $D/r1.fq $D/r2.fq > /dev/null 2> $D/err
python scripts/bench_jobs.py $D/jobs.*.bin > gpurun_out/jobs_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:ksw_fast -s 2 -c 1 -o gpurun_out/prof_jobs_pe150 -f python scripts/bench_jobs.py $D/jobs.*.bin > gpurun_out/jobs_ncu.log 2>&1
tail -4 gpurun_out/jobs_plain.log
rm -rf $D
