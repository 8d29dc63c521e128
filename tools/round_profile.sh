#!/bin/bash
# One GPU-box pass that produces everything profiles/ is built from:
#   tests, bench line, ncu launch list of the same bench command, one ncu --set full
#   capture of the hot kernels (summary CSV + per-instruction stall CSV), kernel microbenchmarks.
# usage (under gpurun): bash tools/round_profile.sh <tag>
tag=${1:-r1}
out=gpurun_out
mkdir -p $out
python -m pytest tests -m gpu -q 2>&1 | tail -3 > $out/${tag}_pytest.log
python bench.py --steps 20 --warmup 5 > $out/${tag}_bench.json 2> $out/${tag}_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
    --log-file $out/${tag}_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline \
    > $out/${tag}_ncu_bench.log 2>&1
ncu --set full --clock-control none --import-source on \
    -k regex:'crf_|logz_|rnn_ws|indices_kernel' -c 14 -o $out/${tag}_full \
    python tools/profile_target.py crf logz rnn > $out/${tag}_full.log 2>&1
ncu -i $out/${tag}_full.ncu-rep --page raw --csv > $out/${tag}_full_raw.csv 2>/dev/null
ncu -i $out/${tag}_full.ncu-rep --page source --csv > $out/${tag}_full_source.csv 2>/dev/null
rm -f $out/${tag}_full.ncu-rep
rm -f $out/microbench.jsonl
python tools/microbench.py crf rnn sweep > $out/${tag}_microbench.log 2>&1
cp $out/microbench.jsonl $out/${tag}_microbench.jsonl
cat $out/${tag}_pytest.log
head -c 400 $out/${tag}_bench.json
