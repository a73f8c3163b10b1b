#!/bin/bash
# Round-end measurement on one B200 (run through gpurun): GPU tests, an extra level-1 fuzz with a fresh seed, the bench
# lines (headline, level 1 on text and random) and the ncu launch list of the headline step.  Outputs in gpurun_out/.
tag=${1:-final}
o=gpurun_out
python -m pytest tests -m gpu -x -q > $o/${tag}_gputests.log 2>&1; echo "pytest rc=$?" >> $o/${tag}_gputests.log
tail -n 3 $o/${tag}_gputests.log
L1_SEED=77 timeout 120 python tools/gpu_l1_check.py 45 0 > $o/${tag}_l1fuzz.log 2>&1; tail -n 1 $o/${tag}_l1fuzz.log
python bench.py > $o/${tag}_bench_text_l2.json 2> $o/${tag}_bench_text_l2.err; echo "bench rc=$?"; cut -c1-400 $o/${tag}_bench_text_l2.json
python bench.py --level 1 --workload text --no-cpu-baseline > $o/${tag}_bench_text_l1.json 2> $o/${tag}_bench_text_l1.err; echo "bench rc=$?"; cut -c1-200 $o/${tag}_bench_text_l1.json
python bench.py --level 1 --workload random --no-cpu-baseline > $o/${tag}_bench_random_l1.json 2> $o/${tag}_bench_random_l1.err; echo "bench rc=$?"; cut -c1-200 $o/${tag}_bench_random_l1.json
B="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-verify"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $o/${tag}_launches.csv $B > $o/${tag}_ncu_l.log 2>&1; echo "ncu rc=$?"
