"""N>1 host logic (view sharding + one all-reduce over the flat gradient buffer) on CPU with gloo, world_size 2.
The per-view renderer is the float64 oracle here (test infrastructure), so the test runs without a GPU."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from parity_common import PKG, ROOT

import bags_synth as S
import bags_view_parallel as VP


def test_shard_views_partition():
    for n in (1, 5, 8, 32):
        for world in (1, 2, 4, 8):
            seen = []
            for r in range(world):
                seen += VP.shard_views(n, r, world)
            assert sorted(seen) == list(range(n))
    assert VP.shard_views(8, 1, 2) == [1, 3, 5, 7]
    with pytest.raises(ValueError):
        VP.shard_views(8, 2, 2)


def test_flat_params_share_storage_and_accumulate():
    cam = S.make_camera(32, 32)
    gs = S.make_gaussians(50, cam)
    fp = VP.FlatGaussianParams(gs.means3D, gs.shs, gs.opacities, gs.scales, gs.rotations)
    assert fp.flat.numel() == 50 * VP.FLOATS_PER_GAUSSIAN
    assert torch.equal(fp["shs"].detach(), gs.shs)
    loss = (fp["means3D"] ** 2).sum() + fp["opacities"].sum()
    loss.backward()
    loss2 = (fp["means3D"] ** 2).sum()
    loss2.backward()
    assert torch.allclose(fp.flat_grad[:150].view(50, 3), 4 * gs.means3D)      # accumulated in place, twice
    assert fp["means3D"].grad.data_ptr() == fp.flat_grad.data_ptr()
    off = 50 * (3 + 48)
    assert torch.allclose(fp.flat_grad[off:off + 50], torch.ones(50))
    fp.zero_grad()
    assert float(fp.flat_grad.abs().sum()) == 0.0


def _oracle_view_loss(params, view):
    sys.path.insert(0, ROOT)
    from oracle import raster_oracle as O
    cam, up = view

    class F(torch.autograd.Function):
        @staticmethod
        def forward(ctx, means3D, shs, opac, scales, rots):
            sc = O.Scene(means3D, opac, cam.viewmatrix, cam.projmatrix, cam.campos, cam.W, cam.H, cam.tanfovx,
                         cam.tanfovy, (0, 0, 0), shs=shs, scales=scales, rotations=rots,
                         shift_factors=torch.zeros(3), sh_degree=3)
            f = O.rasterize_forward(sc)
            ctx.sc, ctx.f = sc, f
            return f["color"].float(), f["depth"].float()[None], f["alpha"].float()[None]

        @staticmethod
        def backward(ctx, gc, gd, ga):
            b = O.rasterize_backward(ctx.sc, ctx.f, gc, gd, ga)
            return (b["dL_dmeans3D"].float(), b["dL_dsh"].float(), b["dL_dopacity"].float(), b["dL_dscales"].float(),
                    b["dL_drotations"].float())

    c, d, a = F.apply(params["means3D"], params["shs"], params["opacities"], params["scales"], params["rotations"])
    return (c * up[0]).sum() + (d * up[1]).sum() + (a * up[2]).sum()


def _make_views(n, W, H):
    return [(S.make_camera(W, H, seed=100 + v), S.make_upstream(H, W, seed=1 + v)) for v in range(n)]


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    W, H, P, NV = 48, 32, 300, 4
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, scale=0.05)
    fp = VP.FlatGaussianParams(gs.means3D, gs.shs, gs.opacities, gs.scales, gs.rotations)
    views = _make_views(NV, W, H)
    VP.run_views(fp, views, _oracle_view_loss, rank=rank, world=world)
    torch.save(fp.flat_grad.clone(), os.path.join(out_dir, f"grad_{rank}.pt"))
    # second iteration with the view -> rank table balanced by measured work (tile instances of each rank's own views,
    # exchanged with one all_gather_object): the reduced gradients do not depend on who rendered what
    sys.path.insert(0, ROOT)
    from oracle import raster_oracle as O
    local = []
    for vi in VP.shard_views(NV, rank, world):
        cam = views[vi][0]
        sc = O.Scene(gs.means3D, gs.opacities, cam.viewmatrix, cam.projmatrix, cam.campos, cam.W, cam.H, cam.tanfovx,
                     cam.tanfovy, (0, 0, 0), shs=gs.shs, scales=gs.scales, rotations=gs.rotations,
                     shift_factors=torch.zeros(3), sh_degree=3)
        local.append(O.rasterize_forward(sc)["num_rendered"])
    costs = VP.gather_view_costs(local, NV, rank, world)
    table = VP.balance_views(costs, world)
    fp.zero_grad()
    VP.run_views(fp, views, _oracle_view_loss, rank=rank, world=world, assignment=table)
    torch.save({"grad": fp.flat_grad.clone(), "table": table, "costs": costs},
               os.path.join(out_dir, f"balanced_{rank}.pt"))
    # third: the sliced exchange (SlicedGradientReducer) -- the buffer is final slice by slice, as the batched backward
    # leaves it; every element must be summed exactly once whatever the slice boundaries are
    g = torch.Generator().manual_seed(7 + rank)
    local = torch.randn(fp.flat_grad.shape, generator=g)
    results = {}
    for bounds in ([0, P], [0, 100, 101, 101, 299, P], [0, 1, P - 1, P]):
        fp.flat_grad.copy_(local)
        red = VP.SlicedGradientReducer(fp)
        for a, b in zip(bounds[:-1], bounds[1:]):
            red.on_slice(a, b - a)
        red.finish()
        results[tuple(bounds)] = fp.flat_grad.clone()
    # the peer-memory reducer needs CUDA IPC: on CPU tensors every rank must agree on the NCCL / gloo fallback
    fp.flat_grad.copy_(local)
    p2p = VP.P2PGradientReducer(fp)
    assert p2p.fallback_reason is not None and "unavailable" in p2p.describe()
    p2p.reduce()
    results["p2p_fallback"] = fp.flat_grad.clone()
    stats = VP.DensificationStats(P, "cpu")
    stats.accum.fill_(1.0 + rank)
    stats.denom.fill_(1.0)
    stats.max_radii2D.fill_(float(10 + rank))
    sg = torch.full((3,), float(rank + 1))
    VP.GradientReducer(fp).reduce_stats(stats, sg)
    torch.save({"local": local, "sliced": results, "accum": stats.accum.clone(), "denom": stats.denom.clone(),
                "max_radii2D": stats.max_radii2D.clone(), "shift_grad": sg}, os.path.join(out_dir, f"sliced_{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_equal_sum_of_single_view_gradients(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    g0 = torch.load(os.path.join(tmp_path, "grad_0.pt"))
    g1 = torch.load(os.path.join(tmp_path, "grad_1.pt"))
    assert torch.equal(g0, g1)                                   # all-reduce leaves identical buffers
    # single-process reference: all four views accumulated on one rank
    W, H, P, NV = 48, 32, 300, 4
    cam0 = S.make_camera(W, H, seed=100)
    gs = S.make_gaussians(P, cam0, scale=0.05)
    fp = VP.FlatGaussianParams(gs.means3D, gs.shs, gs.opacities, gs.scales, gs.rotations)
    VP.run_views(fp, _make_views(NV, W, H), _oracle_view_loss, rank=0, world=1)
    assert float(fp.flat_grad.abs().max()) > 0
    rel = float((g0 - fp.flat_grad).norm() / fp.flat_grad.norm())
    assert rel < 1e-5, rel
    # the balanced assignment: same table and costs on both ranks, every view once, same reduced gradients
    b0 = torch.load(os.path.join(tmp_path, "balanced_0.pt"))
    b1 = torch.load(os.path.join(tmp_path, "balanced_1.pt"))
    assert b0["table"] == b1["table"] and b0["costs"] == b1["costs"]
    assert sorted(v for q in b0["table"] for v in q) == list(range(NV)) and all(c > 0 for c in b0["costs"])
    assert torch.equal(b0["grad"], b1["grad"])
    rel = float((b0["grad"] - fp.flat_grad).norm() / fp.flat_grad.norm())
    assert rel < 1e-5, rel
    # sliced exchange: every slicing gives the plain sum of the two ranks' buffers, bit for bit on both ranks
    s0 = torch.load(os.path.join(tmp_path, "sliced_0.pt"))
    s1 = torch.load(os.path.join(tmp_path, "sliced_1.pt"))
    want = s0["local"] + s1["local"]
    assert len(s0["sliced"]) == 4 and "p2p_fallback" in s0["sliced"]
    for k, v in s0["sliced"].items():
        assert torch.equal(v, want), k
        assert torch.equal(s1["sliced"][k], want), k
    # small collectives: SUM over [accum | accum_abs | denom | shift grad], MAX over max_radii2D
    assert float(s0["accum"][0]) == 3.0 and float(s0["denom"][0]) == 2.0 and float(s0["max_radii2D"][0]) == 11.0
    assert torch.equal(s0["shift_grad"], torch.full((3,), 3.0)) and torch.equal(s1["shift_grad"], s0["shift_grad"])
