#!/bin/bash
# Round 2, GPU visit 19 (final): smoke(), the GPU suite and the default bench line on the final tree.
set -u
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench rc=$?"
python -c "
import json
d=json.load(open('gpurun_out/bench_final.json')); print(round(d['value'],1), 'e2e', round(d['e2e']['value'],1), 'drop_in', round(d['drop_in']['value'],1), round(d['drop_in']['e2e']['value'],1), d['preprocess_binning'], {k:v['ms'] for k,v in d['stages'].items()})"
