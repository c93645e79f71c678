#!/bin/bash
# Round 2, GPU visit 3: full GPU suite (integration, densification, async, oracle-arbitrated pose bars), the other
# BASELINE configs, C4 chain, densify timing, float64 oracle on C2 / blend-heavy C3.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -rs > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
python tests/harness/bench_densify.py 1000000 > gpurun_out/densify.jsonl 2> gpurun_out/densify.err; echo "densify rc=$?"; cat gpurun_out/densify.jsonl
python tests/harness/bench_c4_chain.py > gpurun_out/c4_chain.jsonl 2> gpurun_out/c4_chain.err; echo "c4 rc=$?"; cat gpurun_out/c4_chain.jsonl
python tests/harness/pose_oracle_check.py C2 > gpurun_out/pose_oracle_c2.jsonl 2> gpurun_out/pose_oracle_c2.err; echo "oracle C2 rc=$?"
python tests/harness/pose_oracle_check.py C3 --scale 0.012 > gpurun_out/pose_oracle_c3_s012.jsonl 2> gpurun_out/pose_oracle_c3_s012.err; echo "oracle C3 0.012 rc=$?"
cat gpurun_out/pose_oracle_c2.jsonl gpurun_out/pose_oracle_c3_s012.jsonl
python tests/harness/bench_configs.py --out gpurun_out/configs.jsonl --steps 10 > gpurun_out/configs.log 2>&1; echo "configs rc=$?"
python - <<'PY'
import json
for l in open('gpurun_out/configs.jsonl'):
    d=json.loads(l)
    print(d.get('tag'), d.get('impl','ours'), round(d.get('value',0) or 0,1) if 'value' in d else d.get('ms'), (d.get('preprocess_binning') or {}).get('frac'), d.get('error','')[:200])
PY
ls -la gpurun_out | head -40
