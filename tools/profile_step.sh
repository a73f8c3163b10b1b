#!/bin/bash
# One bench step under ncu: the launch list (durations) and one --set full capture of every kernel of the step.
#   bash tools/profile_step.sh <tag> [extra bench.py args]
tag=$1; shift
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify $*"
$B > gpurun_out/prof_${tag}_plain.json 2> gpurun_out/prof_${tag}_plain.err || { echo "plain run failed"; tail -5 gpurun_out/prof_${tag}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_${tag}.csv $B > gpurun_out/ncu_${tag}_l.log 2>&1
ncu --set full --clock-control none --import-source on -k 'regex:k_candidates|k_lz|k_huffman|k_offsets|k_emit2|k_checksums|k_fixed|k_gather|k_stored' -s 6 -c 6 -f -o gpurun_out/prof_${tag} $B > gpurun_out/ncu_${tag}_f.log 2>&1
tail -3 gpurun_out/ncu_${tag}_f.log
ls -la gpurun_out/prof_${tag}.ncu-rep
