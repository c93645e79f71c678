#!/usr/bin/env python
"""Short workload for ncu captures: N views of the bench scene (C3, 1M Gaussians, 1080p), forward + backward."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch  # noqa: E402
import bags_synth as S  # noqa: E402
import parity_common as pc  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 0.004
impl = sys.argv[3] if len(sys.argv) > 3 else "ours"       # "ref": the unmodified reference build (oracle/_ref)
mod = pc.load_ref() if impl == "ref" else pc.load_ours()
if impl != "ref" and os.environ.get("BAGS_NCU_STAGING"):
    import bags_capi
    bags_capi.lib().bags_set_tuning(7, int(os.environ["BAGS_NCU_STAGING"]))
c = S.CONFIGS["C3"]
cam0 = S.make_camera(c["W"], c["H"], seed=100)
gs = S.make_gaussians(c["P"], cam0, scale=scale)
# one untimed pass first; with `ncu --profile-from-start off` only the second pass is captured
for phase in ("warm", "profiled"):
    if phase == "profiled":
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
    for v in range(n):
        cam = S.make_camera(c["W"], c["H"], seed=100 + v)
        o = pc.run_view(mod, cam, gs, S.make_upstream(c["H"], c["W"], seed=1 + v), torch.device("cuda:0"))
        print(phase, "view", v, "R", o["num_rendered"])
torch.cuda.synchronize()
torch.cuda.profiler.stop()
