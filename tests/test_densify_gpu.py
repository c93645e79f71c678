"""N4 compaction: bags_densify.densify_and_prune against the reference's own GaussianModel.densify_and_prune
(scene/gaussian_model.py:433-447, staged unmodified under oracle/_ref/glue) on identical models, identical optimizer
state and an identically seeded generator: same surviving rows in the same order; gathered values and Adam moments bit
for bit; the two computed fields of split children (xyz, scaling) to fp32 rounding."""
import copy
import os
import types

import numpy as np
import pytest
import torch

import parity_common as pc
from test_reference_glue_gpu import GLUE, load_glue

S = pc.S
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (pc.ref_available() and os.path.isdir(GLUE)),
                                 reason="reference build / staged reference callers (oracle/_ref) not present")]

TRAIN_ARGS = types.SimpleNamespace(percent_dense=0.01, position_lr_init=0.00016, position_lr_final=0.0000016,
                                   position_lr_delay_mult=0.01, position_lr_max_steps=30000, feature_lr=0.0025,
                                   opacity_lr=0.05, scaling_lr=0.005, rotation_lr=0.001)


def build_model(ns, P, seed=0, steps=3):
    """A reference GaussianModel in mid-training state: parameters, Adam moments after a few steps, densification
    statistics."""
    dev = torch.device("cuda")
    cam = S.make_camera(640, 368, seed=100)
    gs = S.make_gaussians(P, cam, seed=seed, scale=0.02)
    g = torch.Generator().manual_seed(11)
    m = ns.GaussianModel(3, 0)
    pcd = ns.BasicPointCloud(points=gs.means3D.numpy(), colors=torch.rand(P, 3, generator=g).numpy(),
                             normals=np.zeros((P, 3), dtype=np.float32))
    m.create_from_pcd(pcd, 1.0)
    m.training_setup(TRAIN_ARGS)
    with torch.no_grad():
        m._scaling.add_((torch.randn(P, 3, generator=g) * 0.8).to(dev))
        m._rotation.copy_(torch.randn(P, 4, generator=g).to(dev))
        m._opacity.copy_((torch.randn(P, 1, generator=g) * 2.5).to(dev))
        m._features_rest.copy_((torch.randn(P, 15, 3, generator=g) * 0.15).to(dev))
    for group in m.optimizer.param_groups:
        group["lr"] = 1e-3
    for _ in range(steps):
        for group in m.optimizer.param_groups:
            p = group["params"][0]
            p.grad = (torch.randn(p.shape, generator=g) * 0.01).to(dev)
        m.optimizer.step()
    m.xyz_gradient_accum = (torch.rand(P, 1, generator=g) * 0.0008).to(dev)
    m.denom = torch.randint(0, 3, (P, 1), generator=g).float().to(dev)           # zeros -> NaN grads -> 0 (:435)
    m.max_radii2D = (torch.rand(P, generator=g) * 40).to(dev)
    return m


def snapshot(m):
    out = {}
    for group in m.optimizer.param_groups:
        p = group["params"][0]
        st = m.optimizer.state.get(p, {})
        out[group["name"]] = (p.detach().clone(), st.get("exp_avg", torch.empty(0)).clone(),
                              st.get("exp_avg_sq", torch.empty(0)).clone(), st.get("step"))
    out["_aux"] = (m.xyz_gradient_accum.clone(), m.denom.clone(), m.max_radii2D.clone())
    return out


@pytest.mark.parametrize("max_screen_size", [None, 20])
def test_densify_and_prune_equals_reference(max_screen_size):
    import bags_densify
    ns = load_glue("ours")
    P = 50_000
    extent = 4.0
    a, b = build_model(ns, P), build_model(ns, P)
    for (ga, gb) in zip(a.optimizer.param_groups, b.optimizer.param_groups):
        assert torch.equal(ga["params"][0], gb["params"][0])
    torch.manual_seed(123)
    a.densify_and_prune(0.0002, 0.005, extent, max_screen_size)                 # the reference's method
    torch.manual_seed(123)
    info = bags_densify.densify_and_prune(b, 0.0002, 0.005, extent, max_screen_size)
    torch.cuda.synchronize()
    sa, sb = snapshot(a), snapshot(b)
    assert info["rows"] == a.get_xyz.shape[0] == b.get_xyz.shape[0]
    assert info["clones"] > 100 and info["children"] > 100 and info["kept"] < P
    n_new = info["clones"] + info["children"]
    for name in ("f_dc", "f_rest", "opacity", "rotation", "xyz", "scaling"):
        pa, ma, va, step_a = sa[name]
        pb, mb, vb, step_b = sb[name]
        assert pa.shape == pb.shape, name
        assert torch.equal(ma, mb) and torch.equal(va, vb), name                 # Adam moments: copies and zeros
        assert float(mb[info["kept"]:].abs().max()) == 0.0
        assert str(step_a) == str(step_b)
        if name in ("xyz", "scaling"):
            head = info["kept"] + info["clones"]
            assert torch.equal(pa[:head], pb[:head]), name                       # kept rows and clones: gathers
            assert torch.allclose(pa[head:], pb[head:], rtol=2e-6, atol=2e-6), name      # children: computed
            assert float((pa[head:] - pb[head:]).abs().max()) <= 2e-6 * (1 + float(pa[head:].abs().max()))
        else:
            assert torch.equal(pa, pb), name                                     # pure gathers: bit for bit
    for x, y in zip(sa["_aux"], sb["_aux"]):
        assert torch.equal(x, y)
    # the model attributes are the optimizer's parameters (what the next render() reads)
    for group in b.optimizer.param_groups:
        assert getattr(b, bags_densify.GROUP_ATTR[group["name"]]) is group["params"][0]
        assert group["params"][0].requires_grad and group["params"][0].is_leaf
    # and the densified model trains on: one more optimizer step on both, same result
    g = torch.Generator().manual_seed(5)
    for group_a, group_b in zip(a.optimizer.param_groups, b.optimizer.param_groups):
        gr = (torch.randn(group_a["params"][0].shape, generator=g) * 0.01).cuda()
        group_a["params"][0].grad = gr.clone()
        group_b["params"][0].grad = gr.clone()
    a.optimizer.step()
    b.optimizer.step()
    for name in ("f_dc", "opacity", "rotation"):
        pa = [gr for gr in a.optimizer.param_groups if gr["name"] == name][0]["params"][0]
        pb = [gr for gr in b.optimizer.param_groups if gr["name"] == name][0]["params"][0]
        assert torch.equal(pa, pb), name
    assert n_new > 0
