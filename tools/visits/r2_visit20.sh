#!/bin/bash
# Round 2, GPU visit 20: every gradient of ours and of the reference vs the float64 oracle at C2 (the rotation gradient
# crossed the 1e-4 ours-vs-reference bar in visit 19), then the full GPU suite with the oracle arbitrating every gradient.
set -u
mkdir -p gpurun_out
python tests/harness/pose_oracle_check.py C2 > gpurun_out/pose_oracle_c2_v20.jsonl 2> gpurun_out/pose_oracle_c2_v20.err; echo "oracle C2 rc=$?"; cut -c1-1200 gpurun_out/pose_oracle_c2_v20.jsonl
python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/pytest_gpu.log
