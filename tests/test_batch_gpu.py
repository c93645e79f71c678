"""B cameras over the same Gaussians as ONE launch set (bags_rasterize_forward_batch / _backward_batch_*,
diff_gaussian_rasterization.GaussianBatchRasterizer; SURVEY.md section 8f row N3 -- the reference's per-face render()
loop of utils/cubemap_utils.py:229,265) against B calls of the single-view API:

* forward: every per-view output and every private state array bit for bit;
* backward: per-view screen-space / pose gradients per view, Gaussian-parameter gradients = sum over the views
  (1e-5 relative L2: the blend backward commits with floating-point atomics, two runs of the SAME code differ by that);
* gradient sinks, sliced backward (+ slice callback), precomputed colour / covariance inputs, B = 1, P = 0, overflow,
  CUDA-graph capture.
"""
import pytest
import torch

import parity_common as pc

S = pc.S
pytestmark = pytest.mark.gpu


def dev():
    return torch.device("cuda:0")


def _cams(B, W, H):
    # different poses AND different fields of view per camera
    return [S.make_camera(W, H, seed=100 + 7 * v, fovx=S.FOVX * (1.0 + 0.04 * v)) for v in range(B)]


def _scene(B=3, P=20000, W=320, H=208, scale=0.02):
    cams = _cams(B, W, H)
    gs = S.make_gaussians(P, cams[0], seed=0, scale=scale)
    ups = [S.make_upstream(H, W, seed=1 + v) for v in range(B)]
    return cams, gs, ups


def _leaves(gs, d, colors_precomp=None, cov3D_precomp=None):
    leaf = lambda t: t.to(d).clone().requires_grad_(True)
    L = dict(means3D=leaf(gs.means3D), opacities=leaf(gs.opacities))
    if colors_precomp is None:
        L["shs"] = leaf(gs.shs)
    else:
        L["colors_precomp"] = leaf(colors_precomp)
    if cov3D_precomp is None:
        L["scales"] = leaf(gs.scales)
        L["rotations"] = leaf(gs.rotations)
    else:
        L["cov3D_precomp"] = leaf(cov3D_precomp)
    return L


def _settings(mod, cam, d, bg, leaves=True):
    mk = (lambda t: t.to(d).clone().requires_grad_(True)) if leaves else (lambda t: t.to(d))
    view, proj, campos = mk(cam.viewmatrix), mk(cam.projmatrix), mk(cam.campos)
    st = mod.GaussianRasterizationSettings(
        image_height=cam.H, image_width=cam.W, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=bg, scale_modifier=1.0,
        viewmatrix=view, projmatrix=proj, intrinsic=cam.intrinsic.to(d), sh_degree=3, campos=campos, prefiltered=False,
        debug=False, debug_iter=0)
    return st, dict(viewmatrix=view, projmatrix=proj, campos=campos)


def run_batch(mod, cams, gs, ups, capacity, shift=(1e-3, -5e-4, 2e-4), bg=(0.1, 0.2, 0.3), backward=True,
              slices=1, on_slice=None, colors_precomp=None, cov3D_precomp=None, sinks=None, sink_overwrite=False):
    d = dev()
    B = len(cams)
    bgt = torch.tensor(bg, dtype=torch.float32, device=d)
    L = _leaves(gs, d, colors_precomp, cov3D_precomp)
    P = L["means3D"].shape[0]
    sets, camleaves = zip(*[_settings(mod, c, d, bgt) for c in cams])
    shift_t = torch.tensor(shift, dtype=torch.float32, device=d).requires_grad_(True)
    m2d = torch.zeros(B, P, 3, device=d, requires_grad=True)
    m2dd = torch.zeros(B, P, 3, device=d, requires_grad=True)
    rast = mod.GaussianBatchRasterizer(list(sets), capacity, backward_slices=slices, on_slice=on_slice,
                                       sink_overwrite=sink_overwrite)
    kw = {k: v for k, v in L.items() if k not in ("means3D", "opacities")}
    if "cov3D_precomp" in kw:
        kw["cov3D_precomp"] = kw.pop("cov3D_precomp")
    ctx = mod.grad_sinks(True) if sinks is not None else torch.enable_grad()
    if sinks is not None:
        for k, buf in sinks.items():
            mod.register_grad_sink(L[k], buf)
    try:
        with ctx:
            color, radii, depth, weights, mean2D = rast(L["means3D"], m2d, m2dd, shift_t, L["opacities"], **kw)
            out = dict(color=color, radii=radii, depth=depth, weights=weights, mean2D=mean2D, rast=rast)
            saved = color.grad_fn.saved_tensors
            out["buffers"] = (saved[7], saved[8], saved[9])
            if backward:
                loss = 0
                for v in range(B):
                    dc, dd, da = [u.to(d) for u in ups[v]]
                    loss = loss + (color[v] * dc).sum() + (depth[v] * dd).sum() + (weights[v] * da).sum()
                loss.backward()
    finally:
        if sinks is not None:
            for k in sinks:
                mod.unregister_grad_sink(L[k])
    torch.cuda.synchronize()
    if backward:
        g = {k: (None if v.grad is None else v.grad.detach().clone()) for k, v in L.items()}
        g["shift_factors"] = shift_t.grad
        g["means2D"] = m2d.grad
        g["means2D_densify"] = m2dd.grad
        for k in ("viewmatrix", "projmatrix", "campos"):
            g[k] = [cl[k].grad for cl in camleaves]
        out["grads"] = g
    return out


def run_singles(mod, cams, gs, ups, shift=(1e-3, -5e-4, 2e-4), bg=(0.1, 0.2, 0.3), **kw):
    return [pc.run_view(mod, c, gs, u, dev(), shift=shift, bg=bg, **kw) for c, u in zip(cams, ups)]


def _state_of_view(mod, out, v, P, cap, H, W):
    import bags_capi
    L = bags_capi.lib()
    gb, bb, ib = int(L.bags_geom_bytes(P)), int(L.bags_binning_bytes(P, cap, W, H)), int(L.bags_image_bytes(W, H))
    geom, binning, img = out["buffers"]
    assert geom.numel() >= (v + 1) * gb and binning.numel() >= (v + 1) * bb and img.numel() >= (v + 1) * ib
    st = mod._C.unpack_state(P, cap, H, W, geom[v * gb:(v + 1) * gb], binning[v * bb:(v + 1) * bb],
                             img[v * ib:(v + 1) * ib])
    d = {}
    for k, t in st.items():
        t = t.cpu()
        if t.dtype == torch.int32 and k != "depth_bits":
            t = t.to(torch.int64) & 0xFFFFFFFF
        d[k] = t
    return d


@pytest.mark.parametrize("B", [1, 3, 5])
def test_batch_forward_is_bit_identical_per_view(B):
    mod = pc.load_ours()
    cams, gs, ups = _scene(B)
    singles = run_singles(mod, cams, gs, ups, backward=False)
    cap = max(s["num_rendered"] for s in singles) + 1000
    out = run_batch(mod, cams, gs, ups, cap, backward=False)
    P, H, W = gs.means3D.shape[0], cams[0].H, cams[0].W
    counts = out["rast"].counts
    for v, s in enumerate(singles):
        R = s["num_rendered"]
        assert R > 0 and int(counts[v]) == R
        for k in ("color", "depth", "weights", "mean2D"):
            assert torch.equal(out[k][v].cpu(), s[k]), (v, k)
        assert torch.equal(out["radii"][v].cpu(), s["radii"]), v
        so, sr = _state_of_view(mod, out, v, P, cap, H, W), s["state"]
        for k in ("tiles_touched", "ranges", "n_contrib", "final_T", "xy", "conic_opacity"):
            assert torch.equal(so[k], sr[k]), (v, k)
        vis = s["radii"] > 0
        assert torch.equal(so["rgb"][vis], sr["rgb"][vis]), v
        assert torch.equal(so["depth_bits"][vis], sr["depth_bits"][vis]), v
        assert torch.equal(so["point_list"][:R], sr["point_list"]), v
        assert torch.equal(so["point_list_keys"][:R], sr["point_list_keys"]), v
    assert out["rast"].check() == max(s["num_rendered"] for s in singles)


def _check_grads(out, singles, tol=1e-5):
    g = out["grads"]
    B = len(singles)
    for k in ("means3D", "shs", "opacities", "scales", "rotations", "colors_precomp", "cov3D_precomp",
              "shift_factors"):
        if k not in g or g[k] is None:
            assert all(s["grads"].get(k) is None for s in singles), k
            continue
        want = sum(s["grads"][k].double() for s in singles)
        assert pc.rel_l2(g[k].cpu(), want) <= tol, (k, pc.rel_l2(g[k].cpu(), want))
    for v in range(B):
        for k in ("means2D", "means2D_densify"):
            assert pc.rel_l2(g[k][v].cpu(), singles[v]["grads"][k]) <= tol, (k, v)
        for k in ("projmatrix", "campos"):
            assert pc.rel_l2(g[k][v].cpu(), singles[v]["grads"][k]) <= 20 * tol, (k, v)
        # Q13: ours takes the highest-index visible Gaussian deterministically, in both entry points
        assert pc.rel_l2(g["viewmatrix"][v].cpu(), singles[v]["grads"]["viewmatrix"]) <= 20 * tol, ("viewmatrix", v)


@pytest.mark.parametrize("B,slices", [(1, 1), (3, 1), (4, 3), (5, 7)])
def test_batch_backward_equals_sum_of_single_views(B, slices):
    mod = pc.load_ours()
    cams, gs, ups = _scene(B)
    singles = run_singles(mod, cams, gs, ups)
    cap = max(s["num_rendered"] for s in singles) + 17
    seen = []
    out = run_batch(mod, cams, gs, ups, cap, slices=slices, on_slice=lambda f, c: seen.append((f, c)))
    _check_grads(out, singles)
    P = gs.means3D.shape[0]
    assert len(seen) == slices and seen[0][0] == 0 and sum(c for _, c in seen) == P
    assert all(seen[i][0] + seen[i][1] == seen[i + 1][0] for i in range(len(seen) - 1))


def test_batch_backward_with_precomputed_colour_and_covariance():
    mod = pc.load_ours()
    cams, gs, ups = _scene(3, P=6000)
    g = torch.Generator().manual_seed(5)
    cp = torch.rand(gs.means3D.shape[0], 3, generator=g)
    Rm = torch.stack([S.quat_to_rot(q.double()) for q in gs.rotations]).float()
    Mm = Rm * gs.scales[:, None, :]
    Sig = Mm @ Mm.transpose(1, 2)
    cov = torch.stack([Sig[:, 0, 0], Sig[:, 0, 1], Sig[:, 0, 2], Sig[:, 1, 1], Sig[:, 1, 2], Sig[:, 2, 2]], 1).contiguous()
    singles = run_singles(mod, cams, gs, ups, use_colors_precomp=True, use_cov_precomp=True, colors_precomp=cp,
                          cov3D_precomp=cov)
    cap = max(s["num_rendered"] for s in singles) + 5
    out = run_batch(mod, cams, gs, ups, cap, colors_precomp=cp, cov3D_precomp=cov)
    for v, s in enumerate(singles):
        assert torch.equal(out["color"][v].cpu(), s["color"]), v
    _check_grads(out, singles)


def test_batch_gradient_sinks_accumulate_like_sequential_views():
    """With sinks the batched backward adds buffer + g_0 + g_1 + ... onto the caller's buffers: the same as the
    single-view sink path applied view after view; rows no view sees keep their (non-zero) previous value."""
    mod = pc.load_ours()
    cams, gs, ups = _scene(3)
    d = dev()
    singles = run_singles(mod, cams, gs, ups)
    cap = max(s["num_rendered"] for s in singles)
    gen = torch.Generator().manual_seed(3)
    prior = {k: torch.randn(getattr(gs, k).shape, generator=gen) * 1e-3
             for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    sinks = {k: v.to(d).clone() for k, v in prior.items()}
    out = run_batch(mod, cams, gs, ups, cap, sinks=sinks)
    seen_any = torch.stack([s["radii"] > 0 for s in singles]).any(0)
    for k, buf in sinks.items():
        assert out["grads"][k] is None, k                      # the autograd slot of a sunk input is None
        want = prior[k].double() + sum(s["grads"][k].double() for s in singles)
        assert pc.rel_l2(buf.cpu(), want) <= 1e-5, k
        assert torch.equal(buf.cpu()[~seen_any], prior[k][~seen_any]), k      # untouched rows: bit for bit
    for v in range(3):
        assert pc.rel_l2(out["grads"]["means2D"][v].cpu(), singles[v]["grads"]["means2D"]) <= 1e-5


def test_batch_sink_overwrite_replaces_the_buffer():
    """sink_overwrite: the batch's summed gradients replace whatever the sink buffers held (no zeroing needed)."""
    mod = pc.load_ours()
    cams, gs, ups = _scene(3)
    d = dev()
    singles = run_singles(mod, cams, gs, ups)
    cap = max(s["num_rendered"] for s in singles)
    gen = torch.Generator().manual_seed(4)
    sinks = {k: torch.randn(getattr(gs, k).shape, generator=gen).to(d)          # garbage that must disappear
             for k in ("means3D", "shs", "opacities", "scales", "rotations")}
    out = run_batch(mod, cams, gs, ups, cap, sinks=sinks, sink_overwrite=True, slices=2)
    seen_any = torch.stack([s["radii"] > 0 for s in singles]).any(0)
    for k, buf in sinks.items():
        assert out["grads"][k] is None, k
        want = sum(s["grads"][k].double() for s in singles)
        assert pc.rel_l2(buf.cpu(), want) <= 1e-5, k
        assert float(buf.cpu()[~seen_any].abs().max()) == 0.0, k               # unseen rows: exact zeros


def test_batch_overflow_reported_and_no_gaussians():
    mod = pc.load_ours()
    cams, gs, ups = _scene(2, P=4000)
    singles = run_singles(mod, cams, gs, ups, backward=False)
    out = run_batch(mod, cams, gs, ups, min(s["num_rendered"] for s in singles) // 2)
    with pytest.raises(RuntimeError, match="overflowed"):
        out["rast"].check()
    assert torch.isfinite(out["color"]).all()
    empty = gs._replace(means3D=gs.means3D[:0], scales=gs.scales[:0], rotations=gs.rotations[:0],
                        opacities=gs.opacities[:0], shs=gs.shs[:0])
    out = run_batch(mod, cams, empty, ups, 16)
    assert float(out["color"].abs().max()) == 0.0 and out["color"].shape[0] == 2


def test_batch_step_is_graph_capturable():
    """Forward + backward of a batch is a pure enqueue: captured once, replayed, same gradients as eager."""
    mod = pc.load_ours()
    d = dev()
    cams, gs, ups = _scene(3, P=8000)
    B = 3
    bgt = torch.zeros(3, device=d)
    L = _leaves(gs, d)
    sets = [_settings(mod, c, d, bgt, leaves=False)[0] for c in cams]
    P = L["means3D"].shape[0]
    m2d = torch.zeros(B, P, 3, device=d, requires_grad=True)
    m2dd = torch.zeros(B, P, 3, device=d, requires_grad=True)
    shift_t = torch.zeros(3, device=d)
    upc = torch.stack([u[0] for u in ups]).to(d)
    cap = 400000
    rast = mod.GaussianBatchRasterizer(sets, cap)
    bufs = {k: torch.zeros_like(v) for k, v in L.items()}
    for k, v in L.items():
        mod.register_grad_sink(v, bufs[k])

    def body():
        with mod.grad_sinks(True):
            color, radii, depth, weights, mean2D = rast(L["means3D"], m2d, m2dd, shift_t, L["opacities"], shs=L["shs"],
                                                        scales=L["scales"], rotations=L["rotations"])
            (color * upc).sum().backward()

    try:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            body()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        eager = {k: b.clone() for k, b in bufs.items()}
        m2d.grad = None
        m2dd.grad = None
        g = torch.cuda.CUDAGraph()
        for b in bufs.values():
            b.zero_()
        with torch.cuda.graph(g):
            body()
        for rep in range(2):
            for b in bufs.values():
                b.zero_()
            g.replay()
            torch.cuda.synchronize()
            rast.check()
            for k in bufs:
                assert pc.rel_l2(bufs[k], eager[k]) <= 1e-5, (rep, k)
    finally:
        for v in L.values():
            mod.unregister_grad_sink(v)
