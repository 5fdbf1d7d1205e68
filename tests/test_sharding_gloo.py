"""CPU, world_size 2, gloo: the N>1 merge logic of grid sharding (partial sums all-reduced, per-cell arg-max merged
with the lowest-global-index tie rule) reproduces the unsharded oracle reductions."""
import os

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import coverage as ocov
from tests import synth


def _local_partials(xy, f, mu, var, vor, lo, hi):
    mem = ocov.membership(vor, xy[lo:hi])
    A = mem.shape[0]
    cent = np.zeros((A, 4))
    lossp = np.zeros((A, 2))
    aval = np.full(A, -np.inf)
    aidx = np.full(A, -1, dtype=np.int64)
    for i, m in enumerate(mem):
        pts, w = xy[lo:hi][m], mu[lo:hi][m]
        cent[i] = [w.sum(), (w * pts[:, 0]).sum(), (w * pts[:, 1]).sum(), m.sum()]
        lossp[i] = [(((pts - vor.filtered_points[i]) ** 2).sum(axis=1) * f[lo:hi][m]).sum(), m.sum()]
        if m.any():
            ids = np.nonzero(m)[0]
            j = int(np.argmax(var[lo:hi][m]))
            aval[i], aidx[i] = var[lo:hi][m][j], lo + ids[j]
    return cent, lossp, aval, aidx


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfgp_coverage_b200 import sharding
    xy = synth.grid(31)
    f = synth.truth_function(xy)
    rng = np.random.default_rng(3)
    mu, var = rng.normal(0.3, 0.2, xy.shape[0]), rng.random(xy.shape[0])
    var[[5, 700]] = 2.0                      # the same maximum in both shards: the lower global index must win
    vor = ocov.voronoi_bounded(np.array([[0.5, 0.5], [0.2, 0.8], [0.9, 0.1]]), ocov.bounding_box_of(xy))
    lo, hi = sharding.shard_bounds(xy.shape[0], world, rank)
    cent, lossp, aval, aidx = _local_partials(xy, f, mu, var, vor, lo, hi)
    res = dict(cent=torch.from_numpy(cent), lossp=torch.from_numpy(lossp), amax_val=torch.from_numpy(aval),
               amax_idx=torch.from_numpy(aidx))
    # the packed form the coverage kernels write: [cent | amax_val | amax_idx bits | lossp], merged on the host
    A = cent.shape[0]
    pack = torch.cat([res["cent"].reshape(-1), res["amax_val"], res["amax_idx"].view(torch.float64), res["lossp"].reshape(-1),
                      torch.tensor([rank + 1, 0], dtype=torch.int32).view(torch.float64)])      # ... | tie count (int32)
    hosted = sharding.gather_results_to_host({"pack": pack, "pack_shape": (A, A)})
    sharding.allreduce_partials(res)
    if rank == 0:
        torch.save({k: v.clone() for k, v in res.items()}, out)
        torch.save({k: torch.from_numpy(np.ascontiguousarray(v)) for k, v in hosted.items()}, out + ".host")
    dist.destroy_process_group()


def test_grid_sharded_reductions_match_unsharded(tmp_path):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, 29531, out), nprocs=2, join=True)
    res = torch.load(out)
    xy = synth.grid(31)
    f = synth.truth_function(xy)
    rng = np.random.default_rng(3)
    mu, var = rng.normal(0.3, 0.2, xy.shape[0]), rng.random(xy.shape[0])
    var[[5, 700]] = 2.0
    vor = ocov.voronoi_bounded(np.array([[0.5, 0.5], [0.2, 0.8], [0.9, 0.1]]), ocov.bounding_box_of(xy))
    cent, lossp, aval, aidx = _local_partials(xy, f, mu, var, vor, 0, xy.shape[0])
    assert np.allclose(res["cent"].numpy(), cent, rtol=1e-13, atol=1e-13)
    assert np.allclose(res["lossp"].numpy(), lossp, rtol=1e-13, atol=1e-13)
    assert np.array_equal(res["amax_idx"].numpy(), aidx) and np.array_equal(res["amax_val"].numpy(), aval)
    _, _, idx_o = ocov.compute_max_var(vor, np.column_stack((xy, f)), var)
    assert np.array_equal(res["amax_idx"].numpy(), idx_o)
    hosted = torch.load(out + ".host")            # the one-collective host merge gives the same global results
    assert int(hosted.pop("ties")) == 1 + 2       # tie counts of the ranks add up
    assert np.allclose(hosted["cent"].numpy(), cent, rtol=1e-13, atol=1e-13)
    assert np.allclose(hosted["lossp"].numpy(), lossp, rtol=1e-13, atol=1e-13)
    assert np.array_equal(hosted["amax_idx"].numpy(), aidx) and np.array_equal(hosted["amax_val"].numpy(), aval)


def _bcast_worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from mfgp_coverage_b200 import sharding
    bb = np.array([0.0, 1.0, 0.0, 1.0])
    sets = [synth.agents(7, 3), synth.agents(12, 4)]
    parts = sharding.broadcast_partitions(sets, bb, "cpu")
    torch.save([{"seeds": p.seeds.clone(), "off": p.off.clone(), "poly": p.poly.clone(), "areas": torch.from_numpy(p.areas()),
                 "A": p.A, "inside": p.seeds_inside} for p in parts], out + f".{rank}")
    dist.destroy_process_group()


def test_partitions_built_once_and_broadcast(tmp_path):
    """N > 1: rank 0 runs Qhull, every rank ends up with the same packed cells as a locally built partition."""
    from mfgp_coverage_b200 import _coverage as cv
    out = str(tmp_path / "parts.pt")
    mp.spawn(_bcast_worker, args=(2, 29533, out), nprocs=2, join=True)
    bb = np.array([0.0, 1.0, 0.0, 1.0])
    for rank in range(2):
        got = torch.load(out + f".{rank}")
        for g, pts in zip(got, [synth.agents(7, 3), synth.agents(12, 4)]):
            v = cv.BoundedVoronoi(pts, bb)
            seeds, poly, off = v.flat()
            assert g["A"] == len(v) and g["inside"] == v.seeds_inside
            assert np.array_equal(g["seeds"].numpy(), seeds.reshape(-1))
            assert np.array_equal(g["off"].numpy()[:len(v) + 1], off)
            assert np.array_equal(g["poly"].numpy()[:2 * off[-1]], poly.reshape(-1))
            assert np.array_equal(g["areas"].numpy(), v.areas())


def test_shard_bounds_cover_the_grid():
    from mfgp_coverage_b200 import sharding
    for G in (1, 7, 2601, 1 << 20):
        for w in (1, 2, 3, 8):
            b = [sharding.shard_bounds(G, w, r) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == G and all(b[i][1] == b[i + 1][0] for i in range(w - 1))


def test_merge_argmax_ties_and_empties():
    from mfgp_coverage_b200 import sharding
    vals = torch.tensor([[1.0, 3.0, 0.0], [1.0, 2.0, 0.0]], dtype=torch.float64)
    idxs = torch.tensor([[10, 4, -1], [3, 9, -1]])
    v, i = sharding.merge_argmax(vals, idxs)
    assert i.tolist() == [3, 4, -1] and v[:2].tolist() == [1.0, 3.0]
