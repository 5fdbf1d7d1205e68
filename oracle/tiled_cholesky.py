"""TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): numpy restatements of the two algorithms behind mfgp_cholesky /
mfgp_cholesky_solve (mfgp-coverage_b200/csrc/gp_fit.cu), used by tests/ to pin their logic on the CPU.  The reference itself
calls np.linalg.cholesky (gaussian_process.py:254, :529); these functions describe HOW the device reaches the same factor.

1. pair_step_factor: the 64x64 diagonal-block factorisation of potrf_diag_body -- two columns per step, 2x2 pivot blocks,
   one reciprocal per step, the inverse carried along, square roots only in the final per-pair scaling.
2. task_order / task_dependencies: the ticket order of chol_dataflow_kernel and what every task waits for; a task must only
   depend on smaller tickets (that is the kernel's deadlock-freedom argument).
"""
import numpy as np


def pair_step_factor(S):
    """Lower Cholesky factor L and W = L^-1 of the SPD matrix S (even order), and LAPACK's `info` (0, or 1 + index of the
    first non-positive pivot; on failure L = W = I as the kernel leaves them).  Follows potrf_diag_body step by step:
    U keeps the UNSCALED columns, M (identity at the start) takes the same eliminations; with P = [[a, b], [b, c]] the pivot
    block of columns (j0, j1):  S -= (U P^-1) U^T  right of the pair,  M -= (U P^-1) [M_j0; M_j1]  below it; afterwards, with
    C = chol(P):  L[:, pair] = U C^-T,  W[pair, :] = C^-1 M[pair, :]."""
    S = np.array(S, dtype=np.float64)
    n = S.shape[0]
    assert n % 2 == 0
    U = np.tril(S)
    M = np.eye(n)
    piv = np.empty((n // 2, 3))
    info = 0
    for k in range(n // 2):
        j0, j1 = 2 * k, 2 * k + 1
        a, b, c = U[j0, j0], U[j1, j0], U[j1, j1]
        det = a * c - b * b
        if not (a > 0.0) or not (det > 0.0):
            if info == 0:
                info = (j1 if a > 0.0 else j0) + 1
            a, b, c, det = 1.0, 0.0, 1.0, 1.0
        piv[k] = a, b, c
        idet = 1.0 / det
        qa, qb, qc = c * idet, -b * idet, a * idet                     # P^-1
        u1, u2 = U[:, j0].copy(), U[:, j1].copy()
        t1 = qa * u1 + qb * u2                                           # T = U P^-1
        t2 = qb * u1 + qc * u2
        rows = np.arange(n) > j1
        cols = np.arange(n) > j1
        U[np.ix_(rows, cols)] -= np.outer(t1[rows], u1[cols]) + np.outer(t2[rows], u2[cols])
        U[:] = np.tril(U)                                                # only the lower triangle is kept
        M[rows, :] -= np.outer(t1[rows], M[j0, :]) + np.outer(t2[rows], M[j1, :])
    if info:
        return np.eye(n), np.eye(n), info
    L = np.empty_like(U)
    W = np.empty_like(M)
    for k in range(n // 2):
        a, b, c = piv[k]
        r1 = 1.0 / np.sqrt(a)
        g = b * r1 * r1
        r2 = 1.0 / np.sqrt(c - g * b)
        j0, j1 = 2 * k, 2 * k + 1
        L[:, j0] = U[:, j0] * r1
        L[:, j1] = (U[:, j1] - g * U[:, j0]) * r2
        W[j0, :] = M[j0, :] * r1
        W[j1, :] = (M[j1, :] - g * M[j0, :]) * r2
    return np.tril(L), np.tril(W), 0


# ---- ticket order of the tile-dataflow kernel ---------------------------------------------------------------------------
# Task kinds (chol_dataflow_kernel): ("chain", d): sub-diagonal tile (d, d-1) + diagonal tile (d, d) + factor/inverse of block d;
# ("L", i, c): tile (i, c) of L, i >= c + 2;  ("Y", c, r): tile (c, r) of Y = L^-1 B.

def two_level_factor(S):
    """The diagonal tile of a chain task (csrc/gp_fit.cu, chol_dataflow_kernel): the pair-step sweep on the two 32x32 halves,
    glued by three small products --  L11, W11 = sweep(S11);  L21 = S21 W11^T;  L22, W22 = sweep(S22 - L21 L21^T);
    W21 = -W22 (L21 W11).  Returns (L, W = L^-1, info) like pair_step_factor; the first failing pivot wins."""
    n = S.shape[0]
    h = n // 2
    L11, W11, info1 = pair_step_factor(S[:h, :h])
    L21 = S[h:, :h] @ np.tril(W11).T
    L22, W22, info2 = pair_step_factor(S[h:, h:] - L21 @ L21.T)
    L = np.zeros_like(S)
    W = np.zeros_like(S)
    L[:h, :h], L[h:, :h], L[h:, h:] = L11, L21, L22
    W[:h, :h], W[h:, h:] = W11, W22
    W[h:, :h] = -np.tril(W22) @ (L21 @ np.tril(W11))
    info = info1 if info1 else (info2 + h if info2 else 0)
    return L, W, info


def chain_place(d, chain_la=5):
    """Block column whose ticket batch holds chain task d >= 1: d // chain_la columns ahead of column d - 1 (its k loop is
    the longest of the column, so it starts early and is done accumulating when W_{d-1} arrives)."""
    return max(0, d - 1 - (d // chain_la if chain_la > 0 else 0))


def task_order(nb, nr, chain_la=5):
    """Tasks in ticket order: chain 0; then per block column c: the chain tasks placed there (chain_place), the tiles (i, c)
    with i >= c+2, the Y tiles of block row c.  chain_la = 0: chain c+1 sits in column c's batch (every task then depends on
    smaller tickets only)."""
    order = [("chain", 0)]
    dnext = 1
    for c in range(nb):
        while dnext < nb and chain_place(dnext, chain_la) == c:
            order.append(("chain", dnext))
            dnext += 1
        order += [("L", i, c) for i in range(c + 2, nb)]
        order += [("Y", c, r) for r in range(nr)]
    return order


def early_tasks(nb, nr, chain_la=5):
    """Tasks that wait on a LATER ticket (possible only for chain tasks drawn early) and, per such task, how many tickets
    lie between it and the last ticket it depends on.  Progress argument: tasks that are not early depend on smaller tickets
    only; a CTA holds a ticket only while resident; so as long as fewer CTAs can hold early tasks than the grid has, some
    resident CTA always holds the smallest unfinished non-early ticket, whose dependencies are finished or running."""
    order = task_order(nb, nr, chain_la)
    pos = {t: k for k, t in enumerate(order)}
    out = {}
    for t in order:
        last = max((pos[producer_of(d)] for d in task_dependencies(t)), default=-1)
        if last > pos[t]:
            out[t] = last - pos[t]
    return out


def producer_of(tile):
    """The task that sets the ready flag of a tile: ("Ltile", i, c) with i > c, ("W", c) (diagonal block c factored and
    inverted), ("Ytile", k, r)."""
    if tile[0] == "W":
        return ("chain", tile[1])
    if tile[0] == "Ytile":
        return ("Y", tile[1], tile[2])
    _, i, c = tile
    return ("chain", i) if i == c + 1 else ("L", i, c)


def task_dependencies(task):
    """Every flag a task waits for, as tiles (see producer_of)."""
    if task[0] == "chain":
        d = task[1]
        if d == 0:
            return []
        deps = [("W", d - 1)]
        for k in range(d - 1):                                            # accumulates L_dk L_(d-1)k^T and L_dk L_dk^T
            deps += [("Ltile", d, k), ("Ltile", d - 1, k)]
        return deps
    if task[0] == "L":
        _, i, c = task
        deps = [("W", c)]
        for k in range(c):
            deps += [("Ltile", i, k), ("Ltile", c, k)]
        return deps
    _, c, r = task
    deps = [("W", c)]
    for k in range(c):
        deps += [("Ltile", c, k), ("Ytile", k, r)]
    return deps


def gram_group_rows(nb, mg):
    """Block-row ranges of the Gram groups (df_group_rows / df_gram_groups in csrc/gp_fit.cu): (nb - mg) // mg groups of mg
    rows, then the remaining mg .. 2 mg - 1 rows in groups that halve down to single rows (..., 4, 2, 1, 1): the work left
    behind the last diagonal block is one two-slab task per tile."""
    m_full = (nb - mg) // mg if nb >= 2 * mg else 0
    out = [(g * mg, (g + 1) * mg) for g in range(m_full)]
    kb, left = m_full * mg, nb - m_full * mg
    while left > 0:
        sz = (left + 1) // 2 if left > 1 else left
        out.append((kb, kb + sz))
        kb, left = kb + sz, left - sz
    return out


def gram_tasks(nb, nr, mg):
    """Gram tickets of mfgp_cholesky_solve_gram in ticket order: group-major, then the lower tiles (ti, tj) of M = Y^T Y row by
    row.  Task (g, ti, tj) adds the block rows of group g (gram_group_rows) of Y to the tile, in place."""
    tiles = [(ti, tj) for ti in range(nr) for tj in range(ti + 1)]
    return [("M", g, ti, tj) for g in range(len(gram_group_rows(nb, mg))) for (ti, tj) in tiles]


def gram_dependencies(task, nb, mg):
    """What a Gram task waits for: the Y tiles of its block rows in both tile columns, and the same tile of the group before."""
    _, g, ti, tj = task
    kb, ke = gram_group_rows(nb, mg)[g]
    deps = []
    for k in range(kb, ke):
        deps += [("Y", k, ti), ("Y", k, tj)]
    if g > 0:
        deps.append(("M", g - 1, ti, tj))
    return deps


def simulate_two_queues(nb, nr, mg, m_lead, ncta, rng, chain_la=5, max_steps=None):
    """Event simulation of the kernel's two-queue policy (chol_dataflow_kernel, thread 0 of every CTA) under an ADVERSARIAL
    interleaving: at every step ONE randomly chosen CTA acts -- it finishes its task if the task's dependencies are done,
    otherwise it idles; a free CTA draws by the kernel's rule:
      * a Gram ticket may be CLAIMED when the next one is runnable and the critical queue is >= m_lead columns ahead of the
        chain (or empty); runnable = critical queue empty, or a critical ticket of a column behind the group's last row was
        drawn AND the chain is two diagonal blocks past that row;
      * a claimed ticket that is not runnable (lost race at a group boundary: modelled by claiming WITHOUT the check with
        probability 1/4) is held while the CTA takes critical tickets.
    Returns the completion order; raises AssertionError on deadlock (no CTA can act and work is left)."""
    crit = task_order(nb, nr, chain_la)
    col_of = {}
    c = 0
    dnext = 1
    col_of[("chain", 0)] = 0
    for cc in range(nb):
        while dnext < nb and chain_place(dnext, chain_la) == cc:
            col_of[("chain", dnext)] = cc
            dnext += 1
        for i in range(cc + 2, nb):
            col_of[("L", i, cc)] = cc
        for r in range(nr):
            col_of[("Y", cc, r)] = cc
    gram = gram_tasks(nb, nr, mg)
    deps = {t: [producer_of(d) for d in task_dependencies(t)] for t in crit}
    deps.update({t: gram_dependencies(t, nb, mg) for t in gram})
    done, order = set(), []
    nc = nm = 0                 # ticket counters
    front = drawn = 0           # ctrl[3], ctrl[4]
    held = [None] * ncta        # deferred Gram ticket per CTA
    running = [None] * ncta
    total = len(crit) + len(gram)
    steps = 0
    limit = max_steps or 400 * total + 1000
    idle_streak = 0

    rows = gram_group_rows(nb, mg)

    def runnable(t):
        last_row = rows[t[1]][1] - 1
        return nc >= len(crit) or (drawn > last_row and last_row + 2 <= front)

    while len(done) < total:
        steps += 1
        assert steps < limit, "no progress"
        k = int(rng.integers(ncta))
        if running[k] is not None:
            t = running[k]
            if all(d in done for d in deps[t]):
                done.add(t)
                order.append(t)
                running[k] = None
                if t[0] == "chain":
                    front = max(front, t[1] + 1)
                idle_streak = 0
            else:
                idle_streak += 1
                assert idle_streak < 50 * ncta * 50, ("deadlock", [r for r in running if r][:6])
            continue
        # free CTA: the kernel's selection loop (one pass)
        crit_left = nc < len(crit)
        if held[k] is None and nm < len(gram):
            nxt = gram[nm]
            lost_race = rng.integers(4) == 0
            if (runnable(nxt) or lost_race) and (not crit_left or drawn - front >= m_lead or lost_race):
                held[k] = nxt
                nm += 1
        if held[k] is not None and runnable(held[k]):
            running[k], held[k] = held[k], None
            continue
        if crit_left:
            t = crit[nc]
            nc += 1
            running[k] = t
            drawn = max(drawn, col_of[t])
            continue
    return order


def bordered_inverse_append(W, z, k_cols, k_nn, yc_new):
    """The block-bordered update behind mfgp_batch_step (mfgp-coverage_b200/csrc/batched.cu, batch_append_kernel): q samples
    are appended to a model whose factor is held as W = L^-1 (N x N, lower) and z = W (y - m).
        K_new = [[K, k], [k^T, kk]],   l = W k,   S = kk - l^T l = C C^T,
        L_new = [[L, 0], [l^T, C]],    W_new = [[W, 0], [-C^-1 l^T W, C^-1]],    z_new = [z ; W_new[N:, :] (y - m)].
    k_cols[N, q]: covariance of the old points with the new ones, k_nn[q, q]: covariance among the new points including the
    noise + jitter diagonal, yc_new[N + q]: ALL centred observations (old and new).  Returns (W_new, z_new).  The reference
    refits from scratch (gaussian_process.py:266-268, :540-542); the leading block of the factor does not change."""
    W = np.asarray(W, dtype=np.float64)
    N, q = W.shape[0], k_nn.shape[0]
    l = W @ k_cols                                   # N x q
    C = np.linalg.cholesky(k_nn - l.T @ l)
    Ci = np.linalg.inv(C)
    Wn = np.zeros((N + q, N + q))
    Wn[:N, :N] = W
    Wn[N:, :N] = -Ci @ (l.T @ W)
    Wn[N:, N:] = np.tril(Ci)
    zn = np.concatenate((np.asarray(z, dtype=np.float64).reshape(-1), Wn[N:, :] @ np.asarray(yc_new, dtype=np.float64).reshape(-1)))
    return Wn, zn
