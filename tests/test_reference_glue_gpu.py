"""NS-1: the reference's OWN callers run unchanged on top of the drop-in.

`gaussian_renderer/__init__.py:render()` (5-value unpack at :110), `scene/cameras.py:Camera` (learnable
delta_quaternion / delta_translation / learnable_fovx,y -> world_view_transform, full_proj_transform, camera_center,
:359-381), `scene/gaussian_model.py:GaussianModel` (activations, `create_from_pcd` -> `distCUDA2`, :161-190) and
`utils/loss_utils.py` are staged byte for byte by oracle/build_ref.py under the git-ignored oracle/_ref/glue and
imported here twice: once with `diff_gaussian_rasterization` / `simple_knn` bound to this repo's packages and once bound
to the unmodified reference build (oracle/_ref).  The same synthetic scene goes through render() -> L1 + SSIM ->
backward in both, and everything a training iteration consumes is compared: the image, radii, the screen-space
gradient sinks, every Gaussian-parameter gradient and the camera-pose / field-of-view gradients that --opt_cam and
--opt_intrinsic step on.

Only test-side shims are used (no reference file is edited): `plyfile` and `easydict` are absent from this image and are
only touched by code paths the test does not run (PLY io, the Lie helper's constructor).
"""
import importlib
import math
import os
import sys
import types

import numpy as np
import pytest
import torch

import parity_common as pc

S = pc.S
GLUE = os.path.join(pc.REF_DIR, "glue")
pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (pc.ref_available() and os.path.isdir(GLUE)),
                                 reason="reference build / staged reference callers (oracle/_ref) not present")]

_GLUE_MODULES = ("gaussian_renderer", "scene", "scene.cameras", "scene.gaussian_model", "utils", "utils.graphics_utils",
                 "utils.general_utils", "utils.sh_utils", "utils.system_utils", "utils.rigid_utils", "utils.camera",
                 "utils.loss_utils")


def _shim_missing_packages():
    if "plyfile" not in sys.modules:
        try:
            importlib.import_module("plyfile")
        except ImportError:
            m = types.ModuleType("plyfile")
            m.PlyData = type("PlyData", (), {})
            m.PlyElement = type("PlyElement", (), {})
            sys.modules["plyfile"] = m
    if "easydict" not in sys.modules:
        try:
            importlib.import_module("easydict")
        except ImportError:
            m = types.ModuleType("easydict")

            class EasyDict(dict):
                __getattr__ = dict.get
                __setattr__ = dict.__setitem__
            m.EasyDict = EasyDict
            sys.modules["easydict"] = m


class _ViewGradSwitch:
    """Q13: the reference's dL_dviewmatrix is a data race (backward.cu:351-363) -- (the 8 terms of ONE arbitrary visible
    Gaussian) / P -- which flows on into delta_quaternion / delta_translation.  With `zero` set, a test-side proxy of
    GaussianRasterizer cuts exactly that gradient edge (the view matrix handed to the rasterizer; the copies inside
    full_proj_transform / camera_center keep theirs) for BOTH implementations, so the pose gradients can be compared at
    the normal bar."""
    zero = False


def _rasterizer_proxy(real):
    m = types.ModuleType("diff_gaussian_rasterization")
    m.GaussianRasterizationSettings = real.GaussianRasterizationSettings

    class GaussianRasterizer(real.GaussianRasterizer):
        def __init__(self, raster_settings):
            if _ViewGradSwitch.zero:
                vm = raster_settings.viewmatrix + 0
                vm.register_hook(lambda g: torch.zeros_like(g))
                raster_settings = raster_settings._replace(viewmatrix=vm)
            super().__init__(raster_settings)
    m.GaussianRasterizer = GaussianRasterizer
    return m


def load_glue(which):
    """The reference's render / Camera / GaussianModel / loss functions bound to `which` in {"ours", "ref"}."""
    _shim_missing_packages()
    if which == "ours":
        dgr = pc.load_ours()
        knn = importlib.import_module("simple_knn._C")
        knn_pkg = importlib.import_module("simple_knn")
    else:
        dgr = pc.load_ref()
        d = os.path.join(pc.REF_DIR, "simple_knn")
        spec = importlib.util.spec_from_file_location("ref_simple_knn._C", os.path.join(d, "_C.so"))
        knn = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(knn)
        knn_pkg = types.ModuleType("simple_knn")
        knn_pkg._C = knn
    saved = {k: sys.modules.get(k) for k in ("diff_gaussian_rasterization", "simple_knn", "simple_knn._C") +
             _GLUE_MODULES}
    for k in _GLUE_MODULES:
        sys.modules.pop(k, None)
    sys.modules["diff_gaussian_rasterization"] = _rasterizer_proxy(dgr)
    sys.modules["simple_knn"] = knn_pkg
    sys.modules["simple_knn._C"] = knn
    sys.path.insert(0, GLUE)
    try:
        gr = importlib.import_module("gaussian_renderer")
        cams = importlib.import_module("scene.cameras")
        gm = importlib.import_module("scene.gaussian_model")
        lu = importlib.import_module("utils.loss_utils")
        gu = importlib.import_module("utils.graphics_utils")
        assert os.path.realpath(gr.__file__).startswith(os.path.realpath(GLUE))
        ns = types.SimpleNamespace(render=gr.render, Camera=cams.Camera, GaussianModel=gm.GaussianModel,
                                   l1_loss=lu.l1_loss, ssim=lu.ssim, BasicPointCloud=gu.BasicPointCloud,
                                   distCUDA2=knn.distCUDA2, rasterizer=dgr)
    finally:
        sys.path.remove(GLUE)
        for k in _GLUE_MODULES:
            sys.modules.pop(k, None)
        for k, v in saved.items():
            if v is not None:
                sys.modules[k] = v
            else:
                sys.modules.pop(k, None)
    return ns


def _iteration(ns, P, W, H, hybrid, shift_values, seed=0):
    """One training iteration of train.py's shape on a synthetic scene, through the reference's own callers."""
    from PIL import Image
    dev = torch.device("cuda")
    cam = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam, seed=seed, scale=0.02)
    g = torch.Generator().manual_seed(7)

    # --- GaussianModel exactly as Scene.__init__ builds it: create_from_pcd -> distCUDA2 -> log-scales ----------------
    model = ns.GaussianModel(3, 0)
    pcd = ns.BasicPointCloud(points=gs.means3D.numpy(), colors=torch.rand(P, 3, generator=g).numpy(),
                             normals=np.zeros((P, 3), dtype=np.float32))
    model.create_from_pcd(pcd, 1.0)
    knn_scaling = model._scaling.detach().clone()
    with torch.no_grad():            # a trained-looking state: anisotropic scales, random rotations / opacities / SH
        model._scaling.add_((torch.randn(P, 3, generator=g) * 0.5).to(dev))
        model._rotation.copy_(torch.randn(P, 4, generator=g).to(dev))
        model._opacity.copy_((torch.randn(P, 1, generator=g) * 1.5).to(dev))
        model._features_rest.copy_((torch.randn(P, 15, 3, generator=g) * 0.15).to(dev))
    model.active_sh_degree = 3

    # --- Camera with learnable pose / field of view (scene/cameras.py) -------------------------------------------------
    w2c = torch.linalg.inv(cam.c2w)
    R = w2c[:3, :3].t().numpy().copy()          # the loaders store the transposed rotation (getWorld2View2)
    T = w2c[:3, 3].numpy().copy()
    fovx = 2 * math.atan(cam.tanfovx)
    fovy = 2 * math.atan(cam.tanfovy)
    img = Image.fromarray((torch.rand(H, W, 3, generator=g).numpy() * 255).astype(np.uint8))
    camera = ns.Camera(0, R, T, np.eye(4, dtype=np.float32), fovx, fovy, W / (2 * cam.tanfovx), H / (2 * cam.tanfovy),
                       img, None, img, "synthetic", 0, orig_fov_w=W, orig_fov_h=H,
                       original_image_resolution=(3, H, W), fish_gt_image_resolution=(3, H, W))
    with torch.no_grad():            # away from the initial pose, so that every pose-gradient term is exercised
        camera.delta_quaternion.add_(torch.tensor([0.0, 0.004, -0.003, 0.002], device=dev))
        camera.delta_translation.add_(torch.tensor([[0.01], [-0.02], [0.015]], device=dev))

    pipe = types.SimpleNamespace(debug=False, compute_cov3D_python=False, convert_SHs_python=False)
    bg = torch.tensor([0.1, 0.2, 0.3], device=dev)
    shift = torch.tensor(shift_values, dtype=torch.float32, device=dev, requires_grad=True)
    align = [torch.eye(3, device=dev), torch.tensor([1.0], device=dev)]
    gt = camera.original_image.to(dev)

    pkg = ns.render(camera, model, pipe, bg, 0.0, shift, hybrid=hybrid, iteration=1, global_alignment=align)
    image = pkg["render"]
    Ll1 = ns.l1_loss(image, gt)
    loss = (1.0 - 0.2) * Ll1 + 0.2 * (1.0 - ns.ssim(image, gt))           # train.py:325
    loss.backward()
    torch.cuda.synchronize()
    out = {
        "image": image.detach(), "radii": pkg["radii"], "depth": pkg["depth"].detach(), "loss": loss.detach(),
        "knn_scaling": knn_scaling,
        "viewspace_points.grad": pkg["viewspace_points"].grad, "viewspace_points_densify.grad":
        pkg["viewspace_points_densify"].grad,
        "xyz.grad": model._xyz.grad, "features_dc.grad": model._features_dc.grad,
        "features_rest.grad": model._features_rest.grad, "scaling.grad": model._scaling.grad,
        "rotation.grad": model._rotation.grad, "opacity.grad": model._opacity.grad,
        "delta_quaternion.grad": camera.delta_quaternion.grad, "delta_translation.grad": camera.delta_translation.grad,
        "learnable_fovx.grad": camera.learnable_fovx.grad, "learnable_fovy.grad": camera.learnable_fovy.grad,
        "shift_factors.grad": shift.grad,
    }
    return {k: (v.detach().clone() if isinstance(v, torch.Tensor) else v) for k, v in out.items()}


@pytest.mark.parametrize("hybrid", [False, True])
def test_reference_callers_run_unchanged_on_the_drop_in(hybrid):
    """render() + Camera + GaussianModel + L1/SSIM of the reference, unchanged, on ours vs on the reference build."""
    P, W, H = 60000, 640, 368
    shift = (1e-3, -5e-4, 2e-4)
    results = {}
    for zero in (False, True):
        _ViewGradSwitch.zero = zero
        try:
            for which in ("ours", "ref"):
                results[(which, zero)] = _iteration(load_glue(which), P, W, H, hybrid, shift)
        finally:
            _ViewGradSwitch.zero = False

    o, r = results[("ours", False)], results[("ref", False)]
    # create_from_pcd -> distCUDA2 -> log(sqrt(.)): bit-exact (A15)
    assert torch.equal(o["knn_scaling"], r["knn_scaling"])
    assert torch.equal(o["radii"], r["radii"])
    assert int((o["radii"] > 0).sum()) > P // 3
    assert float((o["image"] - r["image"]).abs().max()) <= 1e-5
    assert float((o["depth"] - r["depth"]).abs().max()) <= 1e-5
    assert abs(float(o["loss"]) - float(r["loss"])) <= 1e-6
    report = {}
    for k in o:
        if k.endswith(".grad"):
            assert o[k] is not None and r[k] is not None, k
            report[k] = pc.rel_l2(o[k], r[k])
    print(f"[glue hybrid={hybrid}] rel-L2 ours vs reference:", {k: f"{v:.2e}" for k, v in report.items()})
    # everything the Gaussian optimizer and the densification consume, the intrinsics and the pupil shift: 1e-4
    for k in ("viewspace_points.grad", "viewspace_points_densify.grad", "xyz.grad", "features_dc.grad",
              "features_rest.grad", "scaling.grad", "rotation.grad", "opacity.grad", "learnable_fovx.grad",
              "learnable_fovy.grad", "shift_factors.grad"):
        assert report[k] <= 1e-4, (k, report[k])
    # pose: with the reference's racy view-matrix edge in place the two can only agree up to that one-Gaussian term
    # (ours takes the highest-index visible Gaussian, the reference whichever thread wrote last) ...
    for k in ("delta_quaternion.grad", "delta_translation.grad"):
        assert report[k] <= 5e-3, (k, report[k])
    # ... and at the normal bar once that edge is cut on both sides
    oz, rz = results[("ours", True)], results[("ref", True)]
    for k in ("delta_quaternion.grad", "delta_translation.grad", "learnable_fovx.grad", "learnable_fovy.grad"):
        e = pc.rel_l2(oz[k], rz[k])
        print(f"[glue hybrid={hybrid}] {k} without the racy dL_dviewmatrix edge: {e:.2e}")
        assert e <= 1e-4, (k, e)
        assert float(oz[k].abs().max()) > 0
