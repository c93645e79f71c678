"""The asynchronous forward (bags_rasterize_forward_async / diff_gaussian_rasterization.AsyncForward) and the
CUDA-graph captured step (bags_view_parallel.CapturedStep) against the blocking forward: same bits forward, same
gradients, overflow reported and memory-safe."""
import pytest
import torch

import parity_common as pc

S = pc.S
pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _scene(P=20000, W=320, H=208, scale=0.02, seed=100):
    cam = S.make_camera(W, H, seed=seed)
    gs = S.make_gaussians(P, cam, seed=0, scale=scale)
    up = S.make_upstream(H, W, seed=1)
    return cam, gs, up


@pytest.mark.parametrize("slack", [0, 1, 4096, 1_000_000])
def test_async_forward_equals_blocking_forward(slack):
    mod = pc.load_ours()
    cam, gs, up = _scene()
    ref = pc.run_view(mod, cam, gs, up, dev())
    R = ref["num_rendered"]
    assert R > 0
    with mod.AsyncForward(R + slack, slots=4) as af:
        out = pc.run_view(mod, cam, gs, up, dev())
    torch.cuda.synchronize()
    assert int(af.counts[0]) == R and af.check() == R
    assert out["num_rendered"] == R + slack           # the backward's R is the capacity of the binning buffer
    for k in ("color", "depth", "weights", "mean2D"):
        assert torch.equal(out[k], ref[k]), k
    assert torch.equal(out["radii"], ref["radii"])
    so, sr = out["state"], ref["state"]
    for k in ("tiles_touched", "ranges", "n_contrib", "final_T"):
        assert torch.equal(so[k], sr[k]), k
    assert torch.equal(so["point_list"][:R], sr["point_list"])
    assert torch.equal(so["point_list_keys"][:R], sr["point_list_keys"])
    for k, g in ref["grads"].items():
        if g is None:
            assert out["grads"][k] is None
            continue
        assert pc.rel_l2(out["grads"][k], g) <= 1e-5, k       # run-to-run atomic-order noise only


def test_async_overflow_is_reported_and_memory_safe():
    """A capacity below the actual count truncates the tile lists: nothing may be written out of bounds (run under
    compute-sanitizer in tools/visits), and check() must raise."""
    mod = pc.load_ours()
    cam, gs, up = _scene()
    R = pc.run_view(mod, cam, gs, up, dev(), backward=False)["num_rendered"]
    for cap in (R // 2, 1, 0):
        with mod.AsyncForward(cap, slots=2) as af:
            out = pc.run_view(mod, cam, gs, up, dev())
        torch.cuda.synchronize()
        assert int(af.counts[0]) == R
        with pytest.raises(RuntimeError, match="overflowed"):
            af.check()
        assert af.suggested_capacity() > R
        assert torch.isfinite(out["color"]).all()


def test_async_forward_with_nothing_visible_and_no_gaussians():
    mod = pc.load_ours()
    cam, gs, up = _scene(P=500)
    pw = (cam.c2w @ torch.tensor([0.0, 0.0, -5.0, 1.0], dtype=torch.float64))[:3].float()      # behind the camera
    behind = gs._replace(means3D=pw.expand(gs.means3D.shape[0], 3).contiguous())
    with mod.AsyncForward(1024, slots=2) as af:
        out = pc.run_view(mod, cam, behind, up, dev())
    torch.cuda.synchronize()
    assert int((out["radii"] > 0).sum()) == 0 and int(af.counts[0]) == 0
    assert float(out["color"].abs().max()) == 0.0


def test_backward_with_zero_gaussians_does_not_raise():
    """P == 0: the reference skips its kernels and returns .mean(0) of empty tensors (NaN pose gradients,
    rasterize_points.cu:274); ours must not fail on a zero-block launch."""
    mod = pc.load_ours()
    d = dev()
    cam = S.make_camera(64, 48)
    leaf = lambda t: t.to(d).clone().requires_grad_(True)
    view, proj, campos = leaf(cam.viewmatrix), leaf(cam.projmatrix), leaf(cam.campos)
    means = torch.zeros(0, 3, device=d, requires_grad=True)
    st = mod.GaussianRasterizationSettings(
        image_height=48, image_width=64, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.ones(3, device=d),
        scale_modifier=1.0, viewmatrix=view, projmatrix=proj, intrinsic=cam.intrinsic.to(d), sh_degree=0,
        campos=campos, prefiltered=False, debug=False, debug_iter=0)
    color, radii, depth, weights, m2d = mod.GaussianRasterizer(st)(
        means3D=means, means2D=torch.zeros_like(means, requires_grad=True),
        means2D_densify=torch.zeros_like(means, requires_grad=True), shift_factors=torch.zeros(3, device=d),
        opacities=torch.zeros(0, 1, device=d), shs=torch.zeros(0, 16, 3, device=d),
        scales=torch.zeros(0, 3, device=d), rotations=torch.zeros(0, 4, device=d))
    assert float(color.abs().max()) == 0.0
    color.sum().backward()
    torch.cuda.synchronize()
    assert means.grad is not None and means.grad.shape == (0, 3)
    assert torch.isnan(proj.grad).all() and torch.isnan(campos.grad).all()


@pytest.mark.parametrize("streams", [1, 2])
def test_captured_step_equals_eager_step(streams):
    """Three views (forward + backward into the flat gradient buffer through the sinks) captured in one CUDA graph:
    every replay reproduces the eager blocking step."""
    import bags_view_parallel as VP
    mod = pc.load_ours()
    d = dev()
    W, H, P = 320, 208, 20000
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, seed=0, scale=0.02)
    cams = [S.make_camera(W, H, seed=100 + v) for v in range(3)]
    ups = [[u.to(d) for u in S.make_upstream(H, W, seed=1 + v)] for v in range(3)]
    params = VP.FlatGaussianParams(gs.means3D.to(d), gs.shs.to(d), gs.opacities.to(d), gs.scales.to(d),
                                   gs.rotations.to(d), device=d)
    params.enable_direct_accumulation(mod)
    shift = torch.zeros(3, device=d, requires_grad=True)
    bg = torch.zeros(3, device=d)
    leaves = [(c.viewmatrix.to(d).requires_grad_(True), c.projmatrix.to(d).requires_grad_(True),
               c.campos.to(d).requires_grad_(True), c.intrinsic.to(d)) for c in cams]
    pipe = VP.ViewStreams(d, depth=streams)
    images = [None] * 3

    def body():
        params.zero_grad()
        with params.accumulating():
            pipe.begin()
            for k, (cam, up, (view, proj, campos, intr)) in enumerate(zip(cams, ups, leaves)):
                with pipe.view(k):
                    st = mod.GaussianRasterizationSettings(
                        image_height=H, image_width=W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg,
                        scale_modifier=1.0, viewmatrix=view, projmatrix=proj, intrinsic=intr, sh_degree=3,
                        campos=campos, prefiltered=False, debug=False, debug_iter=0)
                    m2d = torch.zeros_like(params["means3D"], requires_grad=True)
                    m2dd = torch.zeros_like(params["means3D"], requires_grad=True)
                    color, radii, depth, weights, _ = mod.GaussianRasterizer(st)(
                        means3D=params["means3D"], means2D=m2d, means2D_densify=m2dd, shift_factors=shift,
                        opacities=params["opacities"], shs=params["shs"], scales=params["scales"],
                        rotations=params["rotations"])
                    pipe.before_backward()
                    torch.autograd.backward([color, depth, weights], up)
                    pipe.after_backward()
                    images[k] = color.detach()
            pipe.end()

    def clear():
        for v, p, c, _ in leaves:
            v.grad = p.grad = c.grad = None
        shift.grad = None

    try:
        clear()
        body()                                                   # eager, blocking forward
        torch.cuda.synchronize()
        g_eager = params.flat_grad.clone()
        img_eager = [i.clone() for i in images]
        pose_eager = [torch.cat([v.grad.flatten(), p.grad.flatten(), c.grad]).clone() for v, p, c, _ in leaves]
        clear()
        step = VP.CapturedStep(body, d, mod, capacity=2_000_000, forwards_per_step=3, reset=clear)
        for rep in range(3):
            params.flat_grad.fill_(123.0)                        # the replay must start from its own zero_grad
            step.replay()
            torch.cuda.synchronize()
            assert step.check() > 0
            assert pc.rel_l2(params.flat_grad, g_eager) <= 1e-5
            for a, b in zip(images, img_eager):
                assert torch.equal(a, b)
            pose = [torch.cat([v.grad.flatten(), p.grad.flatten(), c.grad]) for v, p, c, _ in leaves]
            for a, b in zip(pose, pose_eager):
                assert pc.rel_l2(a, b) <= 1e-5
        # in-place update of a static input is seen by the next replay
        with torch.no_grad():
            leaves[0][0].copy_(cams[1].viewmatrix.to(d))
            leaves[0][1].copy_(cams[1].projmatrix.to(d))
            leaves[0][2].copy_(cams[1].campos.to(d))
        step.replay()
        torch.cuda.synchronize()
        assert torch.equal(images[0], img_eager[1])
    finally:
        params.disable_direct_accumulation()
