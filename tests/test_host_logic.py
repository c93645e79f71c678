"""Host-side mirror of the reference interface: names, argument validation and error behaviour (no GPU needed)."""
import inspect
import os
import subprocess
import sys

import pytest
import torch

from parity_common import PKG, ROOT, load_ours, ref_available, load_ref

import bags_synth as S


def test_settings_fields_match_reference_layout():
    m = load_ours()
    assert m.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "intrinsic", "sh_degree", "campos", "prefiltered", "debug", "debug_iter")


def test_public_surface():
    m = load_ours()
    for name in ("GaussianRasterizationSettings", "GaussianRasterizer", "rasterize_gaussians", "compute_relocation"):
        assert hasattr(m, name)
    for name in ("rasterize_gaussians", "rasterize_gaussians_backward", "mark_visible", "compute_relocation"):
        assert hasattr(m._C, name)
    sig = inspect.signature(m.GaussianRasterizer.forward)
    assert list(sig.parameters)[1:] == ["means3D", "means2D", "means2D_densify", "shift_factors", "opacities", "shs",
                                        "colors_precomp", "scales", "rotations", "cov3D_precomp"]
    from simple_knn._C import distCUDA2  # noqa: F401


@pytest.mark.skipif(not ref_available(), reason="reference build (oracle/_ref) not present")
def test_signatures_equal_reference():
    ours, ref = load_ours(), load_ref()
    assert ours.GaussianRasterizationSettings._fields == ref.GaussianRasterizationSettings._fields
    for fn in ("forward", "markVisible"):
        assert list(inspect.signature(getattr(ours.GaussianRasterizer, fn)).parameters) == \
            list(inspect.signature(getattr(ref.GaussianRasterizer, fn)).parameters)
    assert list(inspect.signature(ours.rasterize_gaussians).parameters) == \
        list(inspect.signature(ref.rasterize_gaussians).parameters)
    assert list(inspect.signature(ours.compute_relocation).parameters) == \
        list(inspect.signature(ref.compute_relocation).parameters)


def _settings(m):
    cam = S.make_camera(32, 32)
    return m.GaussianRasterizationSettings(
        image_height=32, image_width=32, tanfovx=cam.tanfovx, tanfovy=cam.tanfovy, bg=torch.zeros(3),
        scale_modifier=1.0, viewmatrix=cam.viewmatrix, projmatrix=cam.projmatrix, intrinsic=cam.intrinsic,
        sh_degree=0, campos=cam.campos, prefiltered=False, debug=False, debug_iter=0)


def test_exactly_one_of_sh_or_colors():
    m = load_ours()
    r = m.GaussianRasterizer(_settings(m))
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(z, z, z, torch.zeros(3), torch.zeros(4, 1), shs=None, colors_precomp=None, scales=z, rotations=torch.zeros(4, 4))
    with pytest.raises(Exception, match="SHs or precomputed colors"):
        r(z, z, z, torch.zeros(3), torch.zeros(4, 1), shs=torch.zeros(4, 16, 3), colors_precomp=z, scales=z,
          rotations=torch.zeros(4, 4))


def test_exactly_one_of_scale_rot_or_cov():
    m = load_ours()
    r = m.GaussianRasterizer(_settings(m))
    z = torch.zeros(4, 3)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(z, z, z, torch.zeros(3), torch.zeros(4, 1), colors_precomp=z)
    with pytest.raises(Exception, match="scale/rotation pair or precomputed 3D covariance"):
        r(z, z, z, torch.zeros(3), torch.zeros(4, 1), colors_precomp=z, scales=z, rotations=torch.zeros(4, 4),
          cov3D_precomp=torch.zeros(4, 6))


def test_cpu_tensors_are_rejected_not_silently_computed():
    """There is no CPU path: CPU inputs must raise, never fall back."""
    m = load_ours()
    r = m.GaussianRasterizer(_settings(m))
    z = torch.zeros(4, 3)
    with pytest.raises(RuntimeError, match="no CPU path"):
        r(z, z, z, torch.zeros(3), torch.zeros(4, 1), colors_precomp=z, scales=z, rotations=torch.zeros(4, 4))
    from simple_knn._C import distCUDA2
    with pytest.raises(RuntimeError, match="no CPU path"):
        distCUDA2(torch.zeros(10, 3))


def test_bad_means_shape():
    m = load_ours()
    with pytest.raises(RuntimeError, match="num_points, 3"):
        m._C.rasterize_gaussians(torch.zeros(3), torch.zeros(4, 2), torch.zeros(3), torch.zeros(0), torch.zeros(4, 1),
                                 torch.zeros(4, 3), torch.zeros(4, 4), 1.0, torch.zeros(0), torch.eye(4), torch.eye(4),
                                 torch.eye(4), 1.0, 1.0, 8, 8, torch.zeros(4, 16, 3), 0, torch.zeros(3), False, False)


def test_product_does_not_import_oracle():
    """The product path must not route through the oracle or the reference build."""
    bad = []
    for dirpath, _, files in os.walk(PKG):
        if "build" in dirpath.split(os.sep)[-1:]:
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                if "raster_oracle" in text or "oracle/_ref" in text or "import oracle" in text or "from oracle" in text:
                    bad.append(os.path.join(dirpath, f))
    assert not bad, bad


def test_synthetic_scene_is_deterministic():
    cam = S.make_camera(64, 48, seed=100)
    a = S.make_gaussians(500, cam, seed=0)
    b = S.make_gaussians(500, cam, seed=0)
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    # camera conventions: campos == inverse(view)[3,:3]; projmatrix = view @ P^T
    inv = torch.linalg.inv(cam.viewmatrix.double())
    assert torch.allclose(inv[3, :3].float(), cam.campos, atol=1e-6)
    V, R = S.estimate_tile_instances(a, cam)
    assert 0 < V <= 500 and R >= V


def test_loss_window_taps_equal_reference_window():
    """utils/loss_utils.py:36-38: the kernels' 11 taps must be the reference's `gauss / gauss.sum()` bit for bit (the
    normaliser is torch's fp32 sum, which equals the correctly rounded sum; one ulp off biases SSIM by 1e-5)."""
    import ctypes
    import math
    import bags_capi
    taps = (ctypes.c_float * 11)()
    assert bags_capi.lib().bags_loss_window(ctypes.cast(taps, ctypes.c_void_p)) == 0
    gauss = torch.Tensor([math.exp(-(x - 11 // 2) ** 2 / float(2 * 1.5 ** 2)) for x in range(11)])
    want = gauss / gauss.sum()
    got = torch.tensor(list(taps), dtype=torch.float32)
    assert torch.equal(got.view(torch.int32), want.view(torch.int32))


def test_balance_views_equal_cardinality_and_lower_imbalance():
    """bags_view_parallel.balance_views: every view exactly once, the same number of views per rank (+-1), never worse
    than round-robin, deterministic."""
    import random
    import bags_view_parallel as VP
    rnd = random.Random(3)
    for world in (1, 2, 3, 4, 8):
        for n in (0, 1, world, 4 * world, 4 * world + 3):
            costs = [rnd.uniform(2.5, 9.0) * 1e6 for _ in range(n)]
            table = VP.balance_views(costs, world)
            assert len(table) == world
            assert sorted(v for q in table for v in q) == list(range(n))
            sizes = [len(q) for q in table]
            assert max(sizes) - min(sizes) <= 1
            assert table == VP.balance_views(costs, world)
            if n >= world:
                load = [sum(costs[v] for v in q) for q in table]
                rr = [sum(costs[v] for v in range(r, n, world)) for r in range(world)]
                assert max(load) <= max(rr) * (1 + 1e-12)
    # the bench scene's 32 cameras (tile instances in millions): round-robin 1.15 -> balanced < 1.01
    r32 = [8.8, 4.18, 3.61, 4.86, 4.42, 4.65, 5.32, 5.76, 5.25, 4.38, 5.46, 5.17, 5.36, 4.9, 5.54, 6.11, 4.32, 6.18,
           6.23, 5.08, 7.06, 4.3, 7.22, 5.81, 5.58, 3.23, 4.77, 4.86, 2.71, 5.48, 5.47, 4.6]
    t = VP.balance_views(r32, 8)
    load = [sum(r32[v] for v in q) for q in t]
    assert all(len(q) == 4 for q in t) and max(load) / (sum(load) / 8) < 1.01
    with pytest.raises(ValueError):
        VP.balance_views([1.0], 0)


def test_bench_helpers():
    """bench.py without a GPU: the SURVEY section 8d byte model, and the one-JSON-line guard (native libraries that
    write to file descriptor 1 must not reach the real stdout)."""
    import textwrap
    sys.path.insert(0, ROOT)
    import bench
    # SURVEY 8d example: C3 with V = 0.8 M, R = 8 M, N = 2.07 M -> 3.44 GB per view, total formula
    P, V, R, N, Tt = 1_000_000, 800_000, 8_000_000, 1920 * 1080, 120 * 68
    ab = bench.algorithmic_bytes(P, V, R, N, Tt)
    assert sum(ab.values()) == 332 * P + (238 + 24 * 16) * V + 312 * R + 56 * N + 8 * Tt
    assert abs(sum(ab.values()) / 1e9 - 3.44) < 0.03
    code = textwrap.dedent(f'''
        import os, sys
        sys.path.insert(0, {ROOT!r})
        import bench
        bench.guard_stdout()
        os.write(1, b"NCCL version banner\\\\n")
        print("python print")
        bench.emit({{"ok": 1}})
    ''')
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert r.stdout == '{"ok": 1}' + chr(10)
    assert "NCCL version banner" in r.stderr and "python print" in r.stderr
    # no CUDA device here: bench.py must say so in its one line and fail, never fall back to a CPU path
    if not torch.cuda.is_available():
        r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")], capture_output=True, text=True)
        assert r.returncode != 0 and r.stdout.count(chr(10)) == 1 and "no CUDA device" in r.stdout


def test_batch_rasterizer_rejects_inconsistent_views():
    """GaussianBatchRasterizer validates its settings on the host (no GPU needed): 1..8 views that share image size,
    background, scale modifier and SH degree; capacity within the 30-bit look-back words."""
    import torch
    import diff_gaussian_rasterization as d
    eye = torch.eye(4)

    def st(h=32, w=48, bg=None, deg=3, mod=1.0):
        return d.GaussianRasterizationSettings(h, w, 0.5, 0.5, torch.zeros(3) if bg is None else bg, mod, eye, eye, eye,
                                               deg, torch.zeros(3), False, False, 0)
    r = d.GaussianBatchRasterizer([st(), st()], 1000, backward_slices=0)
    assert r.backward_slices == 1 and r.capacity == 1000 and r.counts.shape == (2,) and not r.sink_overwrite
    for bad in ([], [st()] * 9, [st(), st(h=64)], [st(), st(deg=2)], [st(), st(mod=2.0)], [st(), st(bg=torch.ones(3))]):
        with pytest.raises(ValueError):
            d.GaussianBatchRasterizer(bad, 1000)
    for cap in (-1, 1 << 30):
        with pytest.raises(ValueError):
            d.GaussianBatchRasterizer([st()], cap)
    with pytest.raises(Exception, match="excatly one"):
        r(torch.zeros(4, 3), torch.zeros(2, 4, 3), torch.zeros(2, 4, 3), torch.zeros(3), torch.zeros(4, 1))


def test_sliced_reducer_ranges_cover_the_flat_buffer_once():
    import torch
    import bags_view_parallel as VP
    P = 37
    z = torch.zeros
    fp = VP.FlatGaussianParams(z(P, 3), z(P, 16, 3), z(P, 1), z(P, 3), z(P, 4))
    red = VP.SlicedGradientReducer(fp)
    fp.flat_grad.zero_()
    for a, b in ((0, 5), (5, 5), (5, 30), (30, 37)):
        for t in red.slice_ranges(a, b - a):
            t += 1.0                      # views into the flat buffer
    assert torch.equal(fp.flat_grad, torch.ones_like(fp.flat_grad))
    red.on_slice(0, P)                    # no process group: a no-op
    red.finish()


def test_bench_argument_defaults():
    sys.path.insert(0, ROOT)
    import bench
    a = bench.parse([])
    assert a.mode == "batch" and a.allreduce == "auto" and not a.overlap and a.gpus == 1 and a.views_per_gpu == 4
    assert bench.TUNE_KNOBS["BLEND_BWD_REDUCE"] == 8
