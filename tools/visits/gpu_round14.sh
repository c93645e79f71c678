#!/bin/bash
# GPU-box visit 14 (session 8): l1_u8 tests after the host-division fix; host-side profile of a view (C1 and C3).
set -u
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q --tb=short -k "l1_u8" ) > gpurun_out/pytest_gpu_sel.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu_sel.log
python tools/cpu_profile.py --config C1 --views 300 > gpurun_out/cpu_profile_c1.txt 2>&1
python tools/cpu_profile.py --config C3 --views 60 > gpurun_out/cpu_profile_c3.txt 2>&1
ls -la gpurun_out
