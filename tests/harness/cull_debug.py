"""Development tool: forward with patch culling on vs off must be bit-identical; prints the worst pixel otherwise."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "bundle-adjusting-gaussian-splatting_b200")); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import bags_capi, bags_synth as S, parity_common as pc
import test_full_size_gpu as T
L = bags_capi.lib()
c = S.CONFIGS["C1"]; dev = torch.device("cuda:0")
cam = S.make_camera(c["W"], c["H"]); gs = S.make_gaussians(c["P"], cam, scale=0.004); up = S.make_upstream(c["H"], c["W"])
res = {}
for cull in (1, 0):
    L.bags_set_tuning(2, cull)
    res[cull] = T.run_on_gpu(pc.load_ours(), cam, gs, up, dev, backward=False)
L.bags_set_tuning(2, 1)
r = T.run_on_gpu(pc.load_ref(), cam, gs, up, dev, backward=False)
for a, b, nm in ((res[1], res[0], "cull vs nocull"), (res[0], r, "nocull vs ref"), (res[1], r, "cull vs ref")):
    d = (a["state"]["final_T"] - b["state"]["final_T"]).abs()
    print(nm, "final_T max", float(d.max()), "n>1e-5", int((d > 1e-5).sum()), "n_contrib mismatch", int((a["state"]["n_contrib"] != b["state"]["n_contrib"]).sum()),
          "color max", float((a["color"] - b["color"]).abs().max()))
