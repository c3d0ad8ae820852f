#!/bin/bash
# One-shot GPU session for the numbers the judge reads: bench line, ncu launch list of the same
# command, one full capture of the fused pass.  Run under gpurun from the repo root:
#   gpurun --timeout 1500 -- scripts/profile_round.sh r01c
tag=${1:-rXX}
out=gpurun_out
python bench.py --steps 3 --warmup 3 > $out/bench_$tag.json 2> $out/bench_$tag.err || { echo "bench failed"; tail -5 $out/bench_$tag.err; exit 1; }
cat $out/bench_$tag.json
python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_ref_$tag.json 2>> $out/bench_$tag.err
cat $out/bench_ref_$tag.json
# launch list of the same command (cold-cache, serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -s 280 -c 60 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_launch_$tag.log 2>&1
# full capture of three fused passes inside the timed step
ncu --set full --clock-control none --import-source on -k 'regex:k_fused_pass|qb_jit_pass' -s 290 -c 3 -f -o $out/${tag}_fused_pass \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/ncu_full_$tag.log 2>&1
ls -la $out/${tag}_*
