"""Multi-GPU sharding of the hot path: one process per GPU, torch.distributed for the plumbing.

Two axes, both taken from the reference's own parallelism:

  * GRID sharding (config c4): the G grid points are split into contiguous slices, one per rank -- the reference's
    experimental `predict_multiproc` slicing (gaussian_process_numba.py:478-503).  The posterior is then
    communication-free.  The training factor is needed by every rank: rank 0 factorises and broadcasts W = L^-1 and
    z (NCCL over NVLink; 134 MB at N = 4096), or every rank factorises redundantly (`replicate_factor=True`, no
    collective at all; chosen by measurement).  The coverage step all-reduces only O(agents) numbers:
    per-cell partial sums (SUM) and the per-cell (max variance, first index) pairs (all-gather + local merge, lowest
    global index wins ties exactly as np.argmax on the unsharded array).
  * RUN sharding (config c5, replicate sweeps): independent simulations, `runner._map_sims` -- no collective.

The merge logic is plain tensor code so that it is exercised on CPU with the gloo backend (tests/test_sharding_gloo.py).
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_bounds(G, world, rank):
    """Contiguous slice [lo, hi) of rank `rank` (sizes differ by at most one, like np.array_split)."""
    base, extra = divmod(G, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def merge_argmax(vals, idxs, k0=0.0, rel=0.0):
    """vals[R, A], idxs[R, A] (global indices, -1 = empty) -> per-cell (max value, LOWEST index attaining it).

    Ranks are folded in rank order with the SAME tie rule the kernels apply inside a shard (csrc/argmax.cuh,
    argmax_combine): two candidates are tied when they differ by at most rel * (k0 - smaller value) -- a tolerance
    relative to the variance reduction -- and a tie goes to the lower global index, carrying the larger value.  With
    rel = 0 this is the plain first-index arg-max.  Without it two mirror-image grid points ~1e-16 apart that land on
    different ranks would be resolved by rounding noise instead of by index (the single-GPU kernel and the reference
    pick the lower index)."""
    R = vals.shape[0]
    bv, bi = vals[0].clone(), idxs[0].clone()
    for r in range(1, R):
        v, i = vals[r], idxs[r]
        a_empty, b_empty = bi < 0, i < 0
        lo = torch.minimum(bv, v)
        tol = torch.clamp(rel * (k0 - lo), min=0.0) if rel > 0.0 else torch.zeros_like(lo)
        d = bv - v
        take_b = (-d > tol)
        tie = ~(d > tol) & ~take_b
        nv = torch.where(take_b, v, torch.where(tie, torch.maximum(bv, v), bv))
        ni = torch.where(take_b, i, torch.where(tie, torch.minimum(bi, i), bi))
        nv = torch.where(b_empty, bv, torch.where(a_empty, v, nv))
        ni = torch.where(b_empty, bi, torch.where(a_empty, i, ni))
        bv, bi = nv, ni
    bv = torch.where(bi < 0, torch.full_like(bv, -float("inf")), bv)
    return bv, bi


def allreduce_partials(res, group=None, amax_k0=0.0, amax_rel=0.0):
    """In-place combination of the outputs of CoverageGrid.assign_reduce across ranks.  (amax_k0, amax_rel): the arg-max
    tie rule the kernels were called with (see merge_argmax)."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return res
    world = dist.get_world_size(group)
    if res.get("cent") is not None:
        dist.all_reduce(res["cent"], op=dist.ReduceOp.SUM, group=group)
    if res.get("lossp") is not None:
        dist.all_reduce(res["lossp"], op=dist.ReduceOp.SUM, group=group)
    if res.get("ties") is not None:
        dist.all_reduce(res["ties"], op=dist.ReduceOp.SUM, group=group)
    if res.get("amax_val") is not None:
        A = res["amax_val"].numel()
        vals = [torch.empty_like(res["amax_val"]) for _ in range(world)]
        idxs = [torch.empty_like(res["amax_idx"]) for _ in range(world)]
        dist.all_gather(vals, res["amax_val"].contiguous(), group=group)
        dist.all_gather(idxs, res["amax_idx"].contiguous(), group=group)
        v, i = merge_argmax(torch.stack(vals).reshape(world, A), torch.stack(idxs).reshape(world, A), amax_k0, amax_rel)
        res["amax_val"].copy_(v)
        res["amax_idx"].copy_(i)
    return res


def merge_argmax_host(vals, idxs, k0=0.0, rel=0.0):
    """numpy twin of merge_argmax for host-side merging: vals[R, A], idxs[R, A] (int64, -1 = empty)."""
    vals = np.asarray(vals, dtype=np.float64)
    idxs = np.asarray(idxs, dtype=np.int64)
    bv, bi = vals[0].copy(), idxs[0].copy()
    for r in range(1, vals.shape[0]):
        v, i = vals[r], idxs[r]
        a_empty, b_empty = bi < 0, i < 0
        with np.errstate(invalid="ignore"):
            lo = np.minimum(bv, v)
            tol = np.maximum(rel * (k0 - lo), 0.0) if rel > 0.0 else np.zeros_like(lo)
            d = bv - v
            take_b = -d > tol
            tie = ~(d > tol) & ~take_b
        nv = np.where(take_b, v, np.where(tie, np.maximum(bv, v), bv))
        ni = np.where(take_b, i, np.where(tie, np.minimum(bi, i), bi))
        bv = np.where(b_empty, bv, np.where(a_empty, v, nv))
        bi = np.where(b_empty, bi, np.where(a_empty, i, ni))
    bv = np.where(bi < 0, -np.inf, bv)
    return bv, bi


def gather_results_to_host(res, group=None, amax_k0=0.0, amax_rel=0.0, extras=None):
    """Global per-cell results of CoverageGrid.assign_reduce on the host with ONE collective and ONE device->host copy:
    the packed result buffers of all ranks are all-gathered, copied home, and combined there -- partial sums added in rank
    order (deterministic), arg-max pairs merged with the lowest-global-index rule.  Every rank gets the same dict of numpy
    arrays (cent, amax_val, amax_idx, lossp).  Cheaper than allreduce_partials (two all-reduces, two all-gathers and a
    handful of small kernels) when the results go to the host anyway, as in the coverage loops.
    `extras`: small float64 device tensors of THIS rank (cell areas, status flags) that ride home in the same copy;
    returned as out["extras"] (list of numpy arrays)."""
    from ._coverage import CoverageGrid
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = CoverageGrid.results_to_host(res)
        if extras:
            out["extras"] = [e.detach().cpu().numpy() for e in extras]
        return out
    world = dist.get_world_size(group)
    Ac, Ap = res["pack_shape"]
    pack = res["pack"]
    if pack.is_cuda:
        allp = torch.empty((world,) + tuple(pack.shape), dtype=pack.dtype, device=pack.device)
        dist.all_gather_into_tensor(allp, pack, group=group)
    else:
        parts = [torch.empty_like(pack) for _ in range(world)]
        dist.all_gather(parts, pack, group=group)
        allp = torch.stack(parts)
    ex_host = None
    if allp.is_cuda:
        flat = torch.cat([allp.reshape(-1)] + [e.reshape(-1) for e in extras]) if extras else allp.reshape(-1)
        h = torch.empty(flat.shape, dtype=flat.dtype, pin_memory=True)
        h.copy_(flat, non_blocking=True)
        torch.cuda.current_stream(allp.device).synchronize()
        a = h.numpy()[:allp.numel()].reshape(tuple(allp.shape))
        if extras:
            ex_host, o = [], allp.numel()
            for e in extras:
                ex_host.append(h.numpy()[o:o + e.numel()])
                o += e.numel()
    else:
        a = allp.numpy()
        if extras:
            ex_host = [e.numpy().reshape(-1) for e in extras]
    out = {"cent": None, "amax_val": None, "amax_idx": None, "lossp": None,
           "ties": int(a[:, 6 * Ac + 2 * Ap:].copy().view(np.int32)[:, 0].sum())}
    if Ac:
        cent = np.zeros((Ac, 4))
        for r in range(world):                                   # rank order: the same sum on every rank, every run
            cent += a[r, :4 * Ac].reshape(Ac, 4)
        out["cent"] = cent
        out["amax_val"], out["amax_idx"] = merge_argmax_host(a[:, 4 * Ac:5 * Ac], a[:, 5 * Ac:6 * Ac].copy().view(np.int64),
                                                             amax_k0, amax_rel)
    if Ap:
        lossp = np.zeros((Ap, 2))
        for r in range(world):
            lossp += a[r, 6 * Ac:6 * Ac + 2 * Ap].reshape(Ap, 2)
        out["lossp"] = lossp
    if ex_host is not None:
        out["extras"] = ex_host
    return out


def broadcast_partitions(seed_sets, bounding_box, device, src=0, group=None):
    """Bounded Voronoi partitions for every rank from ONE host: rank `src` runs Qhull on each seed set and broadcasts the
    packed cells (seeds, areas, offsets, vertices: ~0.1 KB per cell) over NCCL; the other ranks only queue the receive.
    The partitions are global (every rank classifies its own grid shard against the same cells), so building them once
    removes N - 1 redundant Qhull runs per step and, more importantly, the wait for the slowest host of the N.
    Returns a list of PackedPartition (accepted by CoverageGrid.assign_reduce; .areas() for the host finishing)."""
    from . import _coverage as cv
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    heads = [cv.seeds_summary(p, bounding_box) for p in seed_sets]            # (A, seeds_inside): cheap, every rank
    sizes = [cv.partition_doubles(A) for A, _ in heads]
    total = sum(sizes)
    if rank == src:
        stage = torch.empty(total, dtype=torch.float64, pin_memory=torch.device(device).type == "cuda")
        host = stage.numpy()
        o = 0
        for p, n in zip(seed_sets, sizes):
            cv.pack_partition(cv.BoundedVoronoi(p, bounding_box), host[o:o + n])
            o += n
        buf = stage.to(device, non_blocking=True)
    else:
        buf = torch.empty(total, dtype=torch.float64, device=device)
    if world > 1:
        dist.broadcast(buf, src=src, group=group)
    out, o = [], 0
    for (A, inside), n in zip(heads, sizes):
        out.append(cv.PackedPartition(buf[o:o + n], A, inside))
        o += n
    return out


def broadcast_factor(engine, src=0, group=None):
    """Broadcast the fitted factor state (W, z, Tt) of `engine` (a DeviceGP) from rank `src` to all ranks."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    engine.ensure_factor()        # a fused / deferred fit leaves only the diagonal-block inverses in W: complete it first
    n = engine.npad
    for t in (engine.W[:n], engine.z[:n], engine.Tt[:n]):
        dist.broadcast(t, src=src, group=group)


class ShardedGrid:
    """Rank-local slice of the grid plus the collectives that turn rank-local reductions into global ones."""

    def __init__(self, xy_host, f_host=None, group=None):
        from ._coverage import CoverageGrid
        from ._engine import TensorAxes, detect_tensor_grid
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        xy_host = np.asarray(xy_host, dtype=np.float64).reshape(-1, 2)
        self.G_total = xy_host.shape[0]
        self.lo, self.hi = shard_bounds(self.G_total, self.world, self.rank)
        tg = detect_tensor_grid(xy_host)
        axes = TensorAxes(tg[0], tg[1], torch.device("cuda", torch.cuda.current_device())) if tg is not None else None
        self.local = CoverageGrid(xy_host[self.lo:self.hi], None if f_host is None else np.asarray(f_host)[self.lo:self.hi],
                                  base_index=self.lo, axes=axes)

    def assign_reduce(self, *a, **k):
        return allreduce_partials(self.local.assign_reduce(*a, **k), self.group, k.get("amax_k0", 0.0), k.get("amax_rel", 0.0))


class ShardedSim:
    """Grid-sharded twin of simulator._Sim (config c4: the fixed grid split into whole-column slices, one per rank).

    Every rank keeps its slice of the grid resident, factorises the (small) training system itself -- the fused
    Cholesky + forward substitution of the factored posterior; the x expansion covers only the slice's interval, so the
    right-hand-side count shrinks with the number of ranks -- and clips the SAME global Voronoi cells on its own device
    (deterministic: identical on every rank, no broadcast).  One step needs ONE collective: the all-gather of the packed
    per-cell results (6 A_c + 2 A_p + 1 doubles per rank), merged on the host in rank order with the kernels' arg-max tie
    rule.  Tie points (grid points within TIE_TOL of a bisector, counted over all ranks) send every rank to Qhull's
    polygons for that step, exactly like the single-GPU path."""

    def __init__(self, xy_shard, f_shard, axes, base_index, bounding_box, group=None):
        from ._coverage import CoverageGrid
        self.group = group
        self.bounding_box = np.asarray(bounding_box, dtype=np.float64)
        self.grid = CoverageGrid(xy_shard, f_shard, base_index=base_index, axes=axes)
        f64 = dict(dtype=torch.float64, device=self.grid.device)
        self.mu = torch.empty(self.grid.G, **f64)
        self.var = torch.empty(self.grid.G, **f64)
        self._clip = [None, None]

    def step(self, model, positions, centroids_t, voronoi="auto"):
        """(loss, centroids[A,2], argmax grid indices[A], max variance[A]) -- global results, the same on every rank."""
        from . import _coverage as cv
        from .gaussian_process import prior_variance
        bb = self.bounding_box
        eng = model.engine
        eng.lazy_check = eng.defer_fit = True
        model.predict_device(self.grid.xy, self.mu, self.var, grid=self.grid)     # queued first; the host builds the cells meanwhile
        if voronoi == "qhull":
            loss_vor, lloyd_vor = cv.BoundedVoronoi(positions, bb), cv.BoundedVoronoi(centroids_t, bb)
        else:
            loss_vor = cv.HybridVoronoi(positions, bb, reuse=self._clip[0])
            lloyd_vor = cv.HybridVoronoi(centroids_t, bb, reuse=self._clip[1])
            self._clip = [loss_vor, lloyd_vor]
        k0 = prior_variance(model.params())
        kw = dict(w=self.mu, var=self.var, amax_k0=k0, amax_rel=cv.AMAX_REL)
        extras = None
        if voronoi != "qhull" and self.grid.xy.is_cuda and len(loss_vor) and len(lloyd_vor):
            # this rank's cell areas, the clip flags and the Cholesky status ride home in the gathered results' copy
            extras = [lloyd_vor.dev_areas[:len(lloyd_vor)], loss_vor.dev_areas[:len(loss_vor)],
                      torch.cat([eng.info.reshape(1), lloyd_vor.flag.reshape(1), loss_vor.flag.reshape(1)]).to(torch.float64)]
        host = gather_results_to_host(self.grid.assign_reduce(lloyd_vor, loss_vor, **kw), self.group, k0, cv.AMAX_REL, extras)
        if extras is not None:
            ac, ap, st = host["extras"]
            eng.raise_for_info(int(st[0]))
            if st[1] != 0 or st[2] != 0:
                raise RuntimeError("cov_voronoi_clip: polygon capacity exceeded")
            lloyd_vor._areas_host, loss_vor._areas_host = ac.copy(), ap.copy()
        if host["ties"] and voronoi != "qhull":
            loss_vor.qhull(), lloyd_vor.qhull()
            host = gather_results_to_host(self.grid.assign_reduce(lloyd_vor, loss_vor, **kw), self.group, k0, cv.AMAX_REL)
        if extras is None:
            eng.check_factor(force=True)
        loss = cv.loss_from_partials(host["lossp"], loss_vor.areas())
        cent = cv.centroids_from_partials(host["cent"], lloyd_vor.areas(), bb[0], bb[1], bb[2], bb[3])
        return loss, cent, host["amax_idx"], host["amax_val"]
