#!/bin/bash
# Round 2, GPU visit 15: blend backward at 64 registers (8 CTAs / SM); pixels-per-thread variants with the shared-memory
# reduction; smoke(); full GPU suite.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -4 gpurun_out/smoke.log
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
for t in "" "--tune BLEND_BWD_PPT=4" "--tune BLEND_BWD_PPT=1" "--tune BLEND_FWD_PPT=4"; do
  tag=$(echo "x$t" | tr -d ' -=')
  python bench.py --no-cpu-baseline --no-drop-in --no-e2e $t > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench $tag rc=$?"
  python -c "
import json
d=json.load(open('gpurun_out/bench_$tag.json')); print('$tag', round(d['value'],1), {k: v['ms'] for k, v in d['stages'].items() if 'blend' in k})"
done
