#!/bin/bash
# The BASELINE.json configurations other than the contract line (configs[2] and the
# chunk-length sweep E), one bench.py JSON line each -> gpurun_out/bench_configs.jsonl
out=gpurun_out/bench_configs.jsonl
: > $out
python bench.py --steps 10 --warmup 3 --model mGru_cat_mod_flipflop >> $out 2>> gpurun_out/bench_configs.err
python bench.py --steps 10 --warmup 3 --model mGru_flipflop >> $out 2>> gpurun_out/bench_configs.err
for t in 1000 2000 8000; do
  python bench.py --steps 10 --warmup 3 --tsig $t >> $out 2>> gpurun_out/bench_configs.err
done
python - <<'PY'
import json
for l in open('gpurun_out/bench_configs.jsonl'):
    d = json.loads(l)
    print(d['config']['workload'], '| %.1f M samples/s | %.2f ms/step | loss kernels %.3f ms | frac %.4f' % (
        d['value'] / 1e6, d['ms_per_step'], d['roofline']['avg_launch_ms'], d['roofline']['frac']))
PY
