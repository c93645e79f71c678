#!/bin/bash
# Round 2, GPU visit 2: new blend backward (scalar recurrence, raw moments, 12-slot butterfly), cp.async staging variant,
# NS-1 integration test, float64 oracle on the GPU for the pose gradients, blend-heavy scene, ncu of the blend kernels.
set -u
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
Q="--no-e2e --no-drop-in --no-cpu-baseline --min-seconds 1"
python bench.py > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err; echo "ours rc=$?"
for t in BLEND_BWD_PPT=1 BLEND_BWD_PPT=4 BLEND_STAGING=1 BLEND_FWD_PPT=4; do
  python bench.py $Q --tune $t > gpurun_out/bench_tune_$t.json 2>/dev/null; echo "$t rc=$?"
done
python bench.py $Q --scale 0.012 > gpurun_out/bench_ours_s012.json 2> gpurun_out/bench_ours_s012.err; echo "s012 rc=$?"
python bench.py $Q --scale 0.012 --tune BLEND_STAGING=1 > gpurun_out/bench_ours_s012_async.json 2>/dev/null; echo "s012 async rc=$?"
python bench.py $Q --scale 0.012 --tune BLEND_BWD_PPT=4 > gpurun_out/bench_ours_s012_ppt4.json 2>/dev/null; echo "s012 ppt4 rc=$?"
python bench.py --impl reference --no-e2e --no-drop-in --scale 0.012 --steps 5 --repeats 3 --min-seconds 1 > gpurun_out/bench_ref_s012.json 2>/dev/null; echo "ref s012 rc=$?"
python tests/harness/pose_oracle_check.py C1 C3 > gpurun_out/pose_oracle.jsonl 2> gpurun_out/pose_oracle.err; echo "oracle rc=$?"
for st in 0 1; do
  BAGS_NCU_STAGING=$st timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:blend -o gpurun_out/blend_staging$st python tests/harness/ncu_target.py 1 0.004 ours > gpurun_out/ncu_blend$st.out 2>&1; echo "ncu $st rc=$?"
done
BAGS_SMOKE_NO_REF=1 timeout 600 compute-sanitizer --tool initcheck --log-file gpurun_out/sanitizer_initcheck.log python __graft_entry__.py smoke > gpurun_out/sanitizer_initcheck.out 2>&1; echo "initcheck rc=$?"
for f in gpurun_out/bench_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    st=d.get('stages') or {}
    print(sys.argv[1], round(d['value'],1), 'bwd', st.get('blend_bwd',{}).get('ms'), 'fwd', st.get('blend_fwd',{}).get('ms'), 'pb', (d.get('preprocess_binning') or {}).get('ms'), 'R', d['workload_stats']['tile_instances_mean'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
cat gpurun_out/pose_oracle.jsonl
ls -la gpurun_out | head -50
