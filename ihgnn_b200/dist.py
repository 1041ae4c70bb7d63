"""Multi-GPU hypergraph convolution: hyperedge partition + row-sharded node tables (SURVEY.md
section 8e; the reference itself is single-device).

Partition (`PartitionPlan`: torch ops on the rank's own GPU, or on the CPU in the gloo tests):
  * nodes are row-sharded in contiguous equal ranges *within each node type*, so every rank owns
    U/G users, Q/G queries and I/G items and the slot -> node-type hoisting stays valid;
  * a hyperedge lives on the rank that owns its user, so the user row is always local and at
    most two of the three gathers are remote;
  * remote query / item rows referenced by local hyperedges form the rank's halo; local node
    numbering is [own users | own queries, halo queries | own items, halo items].
The graph is static across steps (Dataset.py:91-96 caches it), so send / receive row lists are
planned once.

Per layer and direction there are two exchange steps, adjoint to each other:
  halo_exchange   owners send the boundary rows other ranks reference      (all-to-all)
  halo_reduce     ranks send their partial sums of halo rows back; the owner adds its own
                  partial and the received ones in ascending source rank order and applies Dv^-1
                  (all-to-all + one deterministic segmented reduction = reduce-scatter)
With NVLink peer memory (the default): owners write boundary rows straight into the readers' tables
(`ihg_halo_copy`, every rank sweeping its destinations from a different start), and the reduction kernels
that produce the halo partial sums write them straight into their owners' receive buffers
(`ihg_segment_reduce_routed` / `ihg_two_hop_reduce_routed`: posted NVLink stores under the kernel's own gathers,
the halo work items interleaved across owners so that no NVLink ingress is hit by every rank at once); the owner
then adds its own partial and the received ones in fixed order with a local `ihg_segment_reduce`.
`IHG_ROUTED_REDUCE=0` (and graphs large enough to share one table pair between all call sites) keeps the pull
form: the owner reads the holders' partials over NVLink and sums them in one kernel (`ihg_halo_reduce`).  Over
NCCL the steps are all-to-alls plus the library's gather / segmented-reduce kernels.
Dense weight gradients are all-reduced.  The sharded step is sync-free and is captured per rank in a CUDA graph.

CUDA only (NCCL + symmetric memory).  The planner is exercised on CPU by tests/test_dist_gloo.py.
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch


def _splits(n: int, world: int) -> np.ndarray:
    return np.array([(n * r) // world for r in range(world + 1)], dtype=np.int64)


def _balanced_splits(weight: torch.Tensor, world: int) -> np.ndarray:
    """Boundaries b[0..world] of contiguous ranges with ~equal total weight (prefix-sum cut)."""
    n = int(weight.shape[0])
    cum = torch.cumsum(weight.to(torch.int64), 0)
    total = int(cum[-1]) if n else 0
    if total == 0:
        return _splits(n, world)
    targets = torch.tensor([(total * r) // world for r in range(1, world)], dtype=torch.int64, device=weight.device)
    cuts = (torch.searchsorted(cum, targets, right=False) + 1).cpu().numpy() if world > 1 else np.zeros(0, np.int64)
    b = np.zeros(world + 1, dtype=np.int64)
    b[world] = n
    for r in range(1, world):
        b[r] = min(max(int(cuts[r - 1]), int(b[r - 1])), n)
    return b


def _owner_of(ids: torch.Tensor, bounds: np.ndarray) -> torch.Tensor:
    """Index r of the range [bounds[r], bounds[r+1]) holding each id (last r with bounds[r] <= id)."""
    inner = torch.as_tensor(bounds[1:-1], dtype=torch.int64, device=ids.device)
    return torch.bucketize(ids, inner, right=True)


class PartitionPlan:
    """Index plan of one rank, computed with torch ops on `device` (the rank's GPU in production: a few
    sorts over the rank's own hyperedges, so 100 M-hyperedge logs plan in well under a second; CPU in the
    gloo tests).  No communication: every rank derives, from the same global log, what it needs from the
    others AND what the others need from it.

    Node id spaces:
      global   u, U+q, U+Q+i                                   (Helpers/Graph.py:110-111)
      own      [own users | own queries | own items]           rows this rank owns (n_own)
      local    [own rows | halo rows from rank 0 | halo rows from rank 1 | ...]   (n_local);
               the chunk received from a rank holds its queries (ascending id) then its items,
               i.e. the local table IS the all-to-all receive layout: no unpack pass, and the
               halo partial sums travelling back are a contiguous slice.

    Index arrays are kept as torch tensors on `device` (`plan.t["name"]`); reading `plan.name` returns the
    numpy copy (host-side tests and the CPU replay of the choreography use those)."""

    _TENSORS = ("edge_ids", "i3_local", "row_slot", "send_rows", "reduce_rowptr", "reduce_col", "reduce_entries",
                "vertex_degrees_own", "dv_inv_own")

    def __init__(self, user, query, item, user_count: int, query_count: int, item_count: int,
                 world: int, rank: int, device=None):
        dev = torch.device(device) if device is not None else torch.device("cpu")
        as_ids = lambda a: (a if torch.is_tensor(a) else torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64))).to(
            device=dev, dtype=torch.int64)
        user, query, item = as_ids(user), as_ids(query), as_ids(item)
        self.world, self.rank, self.device = world, rank, dev
        self.U, self.Q, self.I = user_count, query_count, item_count
        r = rank
        # users: contiguous ranges holding ~E / world hyperedges each (a hyperedge lives with its
        # user, so equal user COUNTS leave the ranks up to ~11 % apart on Zipf-skewed logs and every
        # barrier waits for the slowest); queries / items: equal row counts
        deg_u = torch.bincount(user, minlength=user_count)
        self.ub = _balanced_splits(deg_u, world)
        self.qb, self.ib = _splits(query_count, world), _splits(item_count, world)
        ub, qb, ib = (int(self.ub[r]), int(self.ub[r + 1])), (int(self.qb[r]), int(self.qb[r + 1])), \
            (int(self.ib[r]), int(self.ib[r + 1]))
        self.Uo, self.Qo, self.Io = ub[1] - ub[0], qb[1] - qb[0], ib[1] - ib[0]
        self.n_own = self.Uo + self.Qo + self.Io
        self.own_bounds = (self.Uo, self.Uo + self.Qo)
        t = {}

        edge_owner = _owner_of(user, self.ub)
        t["edge_ids"] = torch.nonzero(edge_owner == r).view(-1)           # ascending global hyperedge ids
        eu, eq, ei = user[t["edge_ids"]], query[t["edge_ids"]], item[t["edge_ids"]]
        self.edge_count = int(t["edge_ids"].numel())

        # ---- halo = remote queries / items referenced by local hyperedges; ids ascend with the owner rank
        # (contiguous ownership ranges), so the sorted unique lists are already grouped by source rank
        q_mine, i_mine = (eq >= qb[0]) & (eq < qb[1]), (ei >= ib[0]) & (ei < ib[1])
        q_need, i_need = torch.unique(eq[~q_mine]), torch.unique(ei[~i_mine])
        q_owner, i_owner = _owner_of(q_need, self.qb), _owner_of(i_need, self.ib)
        qc = torch.bincount(q_owner, minlength=world)
        ic = torch.bincount(i_owner, minlength=world)
        recv = qc + ic
        self.recv_counts = recv.cpu().numpy().astype(np.int64)
        self.R = int(self.recv_counts.sum())
        self.n_local = self.n_own + self.R
        chunk0 = self.n_own + torch.cumsum(recv, 0) - recv                # first local row of the chunk from rank s
        qstart, istart = torch.cumsum(qc, 0) - qc, torch.cumsum(ic, 0) - ic
        q_local = chunk0[q_owner] + torch.arange(q_need.numel(), device=dev) - qstart[q_owner]
        i_local = chunk0[i_owner] + qc[i_owner] + torch.arange(i_need.numel(), device=dev) - istart[i_owner]
        slot_own = torch.repeat_interleave(torch.arange(3, device=dev), torch.tensor([self.Uo, self.Qo, self.Io], device=dev))
        slot_halo = torch.repeat_interleave(torch.tensor([1, 2], device=dev).repeat(world),
                                            torch.stack([qc, ic], 1).reshape(-1))
        t["row_slot"] = torch.cat([slot_own, slot_halo]).to(torch.int32)  # node type (= hyperedge slot) of every local row

        # ---- local ids of the hyperedges' nodes
        lu = eu - ub[0]
        lq = self.Uo + (eq - qb[0])
        if q_need.numel():
            pos = torch.searchsorted(q_need, eq).clamp_(max=q_need.numel() - 1)
            lq = torch.where(q_mine, lq, q_local[pos])
        li = self.Uo + self.Qo + (ei - ib[0])
        if i_need.numel():
            pos = torch.searchsorted(i_need, ei).clamp_(max=i_need.numel() - 1)
            li = torch.where(i_mine, li, i_local[pos])
        t["i3_local"] = torch.stack([lu, lq, li], 1) if self.edge_count else torch.zeros((0, 3), dtype=torch.int64, device=dev)

        # ---- what I send: for every other rank d the own rows it references, queries (ascending) then items,
        # i.e. d's request list evaluated here.  One sort of the keys (d, own-layout row) over the hyperedges
        # of OTHER ranks that touch my query / item ranges.
        width = max(self.Qo + self.Io, 1)
        foreign = edge_owner != r
        mq = foreign & (query >= qb[0]) & (query < qb[1])
        mi = foreign & (item >= ib[0]) & (item < ib[1])
        keys = torch.cat([edge_owner[mq] * width + (query[mq] - qb[0]),
                          edge_owner[mi] * width + (self.Qo + item[mi] - ib[0])])
        keys = torch.unique(keys)                                          # sorted by (destination rank, type, id)
        dest = torch.div(keys, width, rounding_mode="floor")
        t["send_rows"] = self.Uo + (keys - dest * width)                   # own (= local) row index of every sent row
        self.send_counts = torch.bincount(dest, minlength=world).cpu().numpy().astype(np.int64)
        self.S = int(self.send_counts.sum())

        # ---- ordered sum of halo_reduce: own row v <- its own partial, then the partials held for it by the
        # other ranks in ascending rank.  `reduce_col` indexes the flat receive layout [S rows, by source rank]
        # (NCCL path); `reduce_entries` = (peer slot, row inside that peer's chunk) for the fused peer-memory pull.
        _, order = torch.sort(t["send_rows"], stable=True)
        t["reduce_col"] = order
        counts = torch.bincount(t["send_rows"], minlength=self.n_own) if self.S else torch.zeros(self.n_own, dtype=torch.int64, device=dev)
        t["reduce_rowptr"] = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), torch.cumsum(counts, 0)])
        send_off = torch.as_tensor(np.concatenate([[0], np.cumsum(self.send_counts)]), dtype=torch.int64, device=dev)
        src_rank = dest[order]
        peer_slot = src_rank - (src_rank > r).to(torch.int64)              # index among the ranks other than me
        t["reduce_entries"] = torch.stack([peer_slot, order - send_off[src_rank]], 1).to(torch.int32)

        # ---- global degrees of the rows I own (Graph.py:112,120 on the whole hypergraph)
        deg = torch.cat([deg_u[ub[0]:ub[1]],
                         torch.bincount(query, minlength=query_count)[qb[0]:qb[1]],
                         torch.bincount(item, minlength=item_count)[ib[0]:ib[1]]]).to(torch.float32)
        deg[deg == 0] = 1e-8
        t["vertex_degrees_own"] = deg
        t["dv_inv_own"] = deg.pow(-1)                                      # GnnLayers.py:187: fp32 reciprocal
        self.t = t

    def __getattr__(self, name):
        t = self.__dict__.get("t")
        if t is not None and name in t:
            return t[name].cpu().numpy()
        raise AttributeError(name)

    def batch_rows(self, users: torch.Tensor, queries: torch.Tensor, items: torch.Tensor):
        """Fixed-shape lookup plan of the batch head (RawGnn.py:128-133 over a row-sharded feature
        matrix): for the 3B requested rows [users | queries | items] (per-type GLOBAL ids, any device)
        returns (local_rows int64 [3B], mine bool [3B]) -- the own-layout row of every id this rank owns,
        and for the others a spread-out dummy own row (they are masked to zero before the all-reduce;
        distinct dummies keep the duplicate chains of the deterministic scatter-add short).
        Sync-free: no data-dependent shapes."""
        r, B, dev = self.rank, int(users.numel()), users.device
        ids = torch.cat([users, queries, items]).to(torch.int64)
        cache = getattr(self, "_batch_cache", None)
        if cache is None or cache[0] != (B, dev):
            mk = lambda a, b_, c: torch.tensor([int(a)] * B + [int(b_)] * B + [int(c)] * B, dtype=torch.int64, device=dev)
            cache = ((B, dev), mk(self.ub[r], self.qb[r], self.ib[r]),
                     mk(self.ub[r + 1], self.qb[r + 1], self.ib[r + 1]), mk(0, self.Uo, self.Uo + self.Qo),
                     torch.arange(3 * B, dtype=torch.int64, device=dev) % max(self.n_own, 1))
            self._batch_cache = cache
        _, lo, hi, base, dummy = cache
        mine = (ids >= lo) & (ids < hi)
        return torch.where(mine, ids - lo + base, dummy), mine

    def own_global_ids(self) -> np.ndarray:
        """Global node ids of the own rows, in own-layout order."""
        r = self.rank
        return np.concatenate([np.arange(self.ub[r], self.ub[r + 1]),
                               self.U + np.arange(self.qb[r], self.qb[r + 1]),
                               self.U + self.Q + np.arange(self.ib[r], self.ib[r + 1])])


def halo_work_order(rows: torch.Tensor, n_own: int, recv_counts, world: int, rank: int) -> torch.Tensor:
    """Processing order of the work items of a local table (`rows[k]` = local row of item k, ascending): the
    own rows first, then the halo rows round-robin over their owners -- starting with rank + 1 -- instead of
    owner by owner.  The routed reductions store a halo row's partial sum straight into its owner's receive
    buffer; in row order every rank would write to rank 0 first, then rank 1, ...: seven writers on one NVLink
    ingress while the others idle.  Returns a permutation of arange(len(rows))."""
    idx = torch.arange(rows.numel(), device=rows.device)
    halo = rows >= n_own
    if not bool(halo.any()):
        return idx
    chunk_end = torch.as_tensor(np.cumsum(np.asarray(recv_counts)) + n_own, dtype=torch.int64, device=rows.device)
    hidx = idx[halo]
    owner = torch.bucketize(rows[halo], chunk_end, right=True)                 # ascending: rows are
    first = torch.searchsorted(owner, torch.arange(world, device=rows.device))
    within = torch.arange(hidx.numel(), device=rows.device) - first[owner]
    order = torch.argsort(within * world + (owner - rank - 1) % world, stable=True)
    return torch.cat([idx[~halo], hidx[order]])


# ----------------------------------------------------------------------------------------
# device side (CUDA + NCCL)
# ----------------------------------------------------------------------------------------
class ShardedHyperGraph:
    """The device-resident local piece of a partitioned hypergraph plus its exchange plan."""

    def __init__(self, plan: PartitionPlan, device, group=None):
        from .graph import CsrPlan, csr_from_keys
        self.plan, self.group = plan, group
        dev = torch.device(device)
        if dev.type != "cuda":
            raise RuntimeError("ihgnn_b200.dist runs on CUDA (NCCL) only; the planner (PartitionPlan) is the host part")
        self.device = dev
        t = lambda name, dt=torch.int64: plan.t[name].to(device=dev, dtype=dt).contiguous()
        self.EdgeCount = plan.edge_count
        self.node_count = plan.n_local
        self.n_own, self.n_local, self.S, self.R = plan.n_own, plan.n_local, plan.S, plan.R
        self.i3 = t("i3_local", torch.int32)
        edge_of = torch.arange(plan.edge_count, device=dev, dtype=torch.int32).repeat_interleave(3)
        rowptr, _perm, col = csr_from_keys(self.i3.reshape(-1), plan.n_local, values=edge_of)
        self.rowptr, self.col = rowptr, col
        self.plan_csr = CsrPlan(rowptr, col)
        self._interleave_halo_work(plan)
        self.row_slot = t("row_slot", torch.int32)     # slot(row) for the per-slot gradient reduce
        self.own_bounds = plan.own_bounds              # node types of the own rows (typed Linear)
        self.dv_inv_own = t("dv_inv_own", torch.float32)
        self.send_rows = t("send_rows")
        self.send_counts = [int(x) for x in plan.send_counts]
        self.recv_counts = [int(x) for x in plan.recv_counts]
        self.reduce_rowptr = t("reduce_rowptr", torch.int32)
        self.reduce_csr = CsrPlan(self.reduce_rowptr, t("reduce_col", torch.int32))
        self.reduce_entries = t("reduce_entries", torch.int32)    # (peer slot, row in that peer's chunk) per reduce_col entry
        # one buffer pair for all call sites when per-site tables (about a dozen per model) would be too large
        self.share_buffers = False
        self.p2p = None                                # set by enable_peer_memory()
        # Dv^-1 of every local row (own + halo): exchanged once, the graph is static
        self.dv_inv_local = HaloExchangeFn.apply(self.dv_inv_own.view(-1, 1).expand(-1, 4).contiguous(), self)[:, 0].contiguous()
        if os.environ.get("IHG_P2P", "1") != "0":
            self.enable_peer_memory()

    def _interleave_halo_work(self, plan: PartitionPlan) -> None:
        """Reorder the work items (row chunks) of the local CSR plan with `halo_work_order`.  The records are
        independent ({begin, end, row, partial slot}), so any order gives the same bits."""
        seg = self.plan_csr.seg
        if plan.world < 3 or seg.shape[0] == 0:
            return
        order = halo_work_order(seg[:, 2].to(torch.int64), plan.n_own, plan.recv_counts, plan.world, plan.rank)
        self.plan_csr.seg = seg[order].contiguous()
        self.plan_csr.struct.seg = self.plan_csr.seg.data_ptr()

    @property
    def two_hop_nbr(self) -> torch.Tensor:
        """Neighbour list of the local CSR for `ihg_two_hop_reduce` (built on first use)."""
        return self.plan_csr.two_hop_nbr(self.i3, row_slot=self.row_slot)

    # ---- NVLink peer memory (torch symmetric memory: VMM allocations mapped into every rank) ----
    def enable_peer_memory(self) -> bool:
        """Switch halo_exchange / halo_reduce from NCCL all-to-all to direct peer-memory copies:
        owners write boundary rows straight into the halo tail of the readers' local tables, and
        pull the halo partial sums straight out of the holders' buffers (one kernel per exchange,
        `ihg_halo_copy`, plus a symmetric-memory barrier).  Returns False (and keeps NCCL) when
        symmetric memory is not available."""
        import torch.distributed as dist
        try:
            import torch.distributed._symmetric_memory as symm_mem
        except ImportError:
            return False
        world, rank = self.plan.world, self.plan.rank
        if world < 2 or world > 16:
            return False
        meta = torch.zeros((world, world + 2), dtype=torch.int64, device=self.device)
        mine = torch.tensor([self.n_own, self.n_local] + self.send_counts, dtype=torch.int64, device=self.device)
        dist.all_gather_into_tensor(meta, mine, group=self.group)
        meta = meta.cpu().numpy()
        n_own_of = meta[:, 0]
        send_of = meta[:, 2:]                                   # send_of[s][d] = rows s sends to d = rows d receives from s
        # row where MY chunk starts in peer d's local table (its halo chunks are ordered by source rank)
        self._peer_row = [int(n_own_of[d] + send_of[:rank, d].sum()) for d in range(world)]
        self._max_local = int(meta[:, 1].max())
        # routed reduce: the partial sums I hold for owner d land in d's receive buffer [S_d rows, by source rank]
        # at the row where MY block starts; my local table keeps d's rows in one contiguous chunk
        self._recv_row = [int(send_of[d, :rank].sum()) for d in range(world)]
        self._max_send = int(send_of.sum(1).max())
        self._chunk0 = [int(self.n_own + sum(self.recv_counts[:d])) for d in range(world + 1)]
        # send_rows in the order the push sweeps its destinations (rank + 1, rank + 2, ... cyclically)
        so = self._send_off = np.concatenate([[0], np.cumsum(self.send_counts)]).astype(np.int64)
        self.send_rows_push = torch.cat([self.send_rows[int(so[d]):int(so[d + 1])]
                                         for d in [(rank + 1 + j) % world for j in range(world - 1)]]).contiguous()
        self.share_buffers = self._max_local * 128 * 4 * 14 > (40 << 30)
        # routed reduce needs a receive buffer per call site; with the single shared table pair of very large
        # graphs the (validated) pull path stays
        self.routed = os.environ.get("IHG_ROUTED_REDUCE", "1") != "0" and not self.share_buffers
        self._symm = symm_mem
        self._bufs = {}
        # probe once (allocation + rendezvous + barrier); every rank must agree on the outcome
        ok = 1
        try:
            probe = symm_mem.empty((1024,), dtype=torch.float32, device=self.device)
            hdl = symm_mem.rendezvous(probe, group=self.group if self.group is not None else dist.group.WORLD)
            hdl.barrier(channel=0)
        except Exception:                                   # noqa: BLE001 - any failure means "keep NCCL"
            ok = 0
        flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)
        self.p2p = True if int(flag.item()) == 1 else None
        return self.p2p is True

    def table_head(self, key, cols: int) -> Optional[torch.Tensor]:
        """The own-row part [n_own, cols] of call site `key`'s local table (peer memory only): a producer
        that writes there saves `halo_exchange` the copy of the own rows."""
        if not self.p2p:
            return None
        return self.peer_buffer(("x", key), cols)[0][:self.n_own]

    def reduce_site(self, key, cols: int) -> Optional["_ReduceSite"]:
        """Routed-reduce buffers of call site `key` (None: routed mode off -> the pull path).  Allocation is a
        collective (symmetric memory): every rank reaches the sites in the same order."""
        if not (self.p2p and self.routed) or key is None:
            return None
        k = ("route", key, cols)
        if k not in self._bufs:
            self._bufs[k] = _ReduceSite(self, key, cols)
        return self._bufs[k]

    def peer_buffer(self, key, cols: int):
        """Persistent symmetric [max n_local, cols] fp32 buffer for call site `key` (allocated and
        rendezvoused on first use -- a collective, so every rank must reach it in the same order)."""
        if self.share_buffers:
            # the exchanges of a step are strictly ordered by barriers, so every call site can use
            # the same two buffers -- one local table, one partial-sum table -- except the projected rows of an
            # order-2/3 layer, which its backward reads again (ShardedFeatureInteractFn / _EdgeInteractFn save them).  At
            # 10^8 hyperedges a table is several GB; per-site buffers would not fit.
            role, site = key
            keep = isinstance(site, tuple) and len(site) == 2 and (site[0] == "fi" or site[1] == "xpp")
            key = (role, site if keep else "shared")
        k = (key, cols)
        if k not in self._bufs:
            buf = self._symm.empty((self._max_local, cols), dtype=torch.float32, device=self.device)
            hdl = self._symm.rendezvous(buf, group=self.group if self.group is not None else torch.distributed.group.WORLD)
            world, rank = self.plan.world, self.plan.rank
            base = [int(p) for p in hdl.buffer_ptrs]
            row_bytes = cols * 4
            import ctypes
            peers = [d for d in range(world) if d != rank]
            n = len(peers)
            chunk = (ctypes.c_void_p * n)(*[base[d] + self._peer_row[d] * row_bytes for d in peers])   # my chunk in d's table
            own = (ctypes.c_void_p * n)(*[buf.data_ptr()] * n)
            off = (ctypes.c_int64 * (n + 1))(*([0] + list(np.cumsum([self.send_counts[d] for d in peers]))))
            # the push sweeps its destinations starting with the next rank (every rank a different one at a time)
            rot = [(rank + 1 + j) % world for j in range(n)]
            push_chunk = (ctypes.c_void_p * n)(*[base[d] + self._peer_row[d] * row_bytes for d in rot])
            push_off = (ctypes.c_int64 * (n + 1))(*([0] + list(np.cumsum([self.send_counts[d] for d in rot]))))
            self._bufs[k] = (buf, hdl, chunk, own, off, n, push_chunk, push_off)
        return self._bufs[k]


class _ReduceSite:
    """Buffers of one reduce call site in routed mode: `own` [n_own, cols] (this rank's partial sums of its own
    rows), `recv` [S, cols] (symmetric: the partials the other ranks hold for my rows, written by THEIR reduction
    kernels, blocks by source rank), and the `route` my reduction kernels write through."""

    def __init__(self, g: "ShardedHyperGraph", key, cols: int):
        from . import functional as F_
        world, rank = g.plan.world, g.plan.rank
        buf = g._symm.empty((max(g._max_send, 1), cols), dtype=torch.float32, device=g.device)
        self.hdl = g._symm.rendezvous(buf, group=g.group if g.group is not None else torch.distributed.group.WORLD)
        self.recv_full, self.recv = buf, buf[:g.S]
        self.own = torch.empty((g.n_own, cols), dtype=torch.float32, device=g.device)
        base = [int(p) for p in self.hdl.buffer_ptrs]
        row_bytes = cols * 4
        starts, bases = [0], [self.own.data_ptr()]
        for d in range(world):
            if d == rank or g.recv_counts[d] == 0:
                continue
            starts.append(g._chunk0[d])
            bases.append(base[d] + g._recv_row[d] * row_bytes)
        starts.append(g.n_local)
        assert all(a <= b for a, b in zip(starts, starts[1:])) and (len(starts) == 2 or starts[1] == g.n_own)
        self.route = F_.OutRoute(starts, bases, cols, keep=(buf, self.own))


def _all_to_all(out: torch.Tensor, inp: torch.Tensor, out_counts, in_counts, group) -> None:
    import torch.distributed as dist
    dist.all_to_all_single(out, inp, output_split_sizes=out_counts, input_split_sizes=in_counts, group=group)


class HaloExchangeFn(torch.autograd.Function):
    """x_own [n_own, d] -> x_local [n_local, d] = [x_own ; rows received from the owners]: pack the
    rows other ranks reference, all-to-all straight into the tail of the local table.
    Backward = halo_reduce without scaling."""

    @staticmethod
    def forward(ctx, x_own, g: "ShardedHyperGraph", key=None):
        ctx.g, ctx.key = g, key
        return _halo_exchange(x_own, g, key)

    @staticmethod
    def backward(ctx, dx_local):
        g, key = ctx.g, ctx.key
        dx_local = dx_local.contiguous()
        if g.p2p and key is not None:              # the incoming gradient is an ordinary tensor: stage it
            d = int(dx_local.shape[1])
            buf = reduce_buffer(g, ("xb", key), d, dx_local)
            buf.copy_(dx_local)
            return _halo_reduce(buf, g, None, ("xb", key)), None, None
        return _halo_reduce(dx_local, g, None), None, None


class ShardedScatterMeanFn(torch.autograd.Function):
    """ef [E_local, d] -> Dv^-1 * (sum over ALL hyperedges containing the row) for the own rows:
    local segmented sum over own + halo rows, then halo_reduce.  Backward: halo_exchange of the
    gradient, then the node -> hyperedge gather-sum with the (exchanged-once) Dv^-1 of the local rows."""

    @staticmethod
    def forward(ctx, ef, g: "ShardedHyperGraph", key=None):
        from . import _lib
        from . import functional as F_
        ctx.g, ctx.key = g, key
        ef = _lib.rows_f32(ef)
        d = int(ef.shape[1])
        site = g.reduce_site(("sm", key) if key is not None else None, d)
        if site is not None:
            F_.segment_reduce_routed(g.plan_csr, ef, d, site.route)
            return _finish_routed(site, g, g.dv_inv_own)
        s_local = F_.segment_reduce(g.plan_csr, ef, d, out=reduce_buffer(g, ("sm", key), d, ef))
        return _halo_reduce(s_local, g, g.dv_inv_own, ("sm", key) if key is not None else None)

    @staticmethod
    def backward(ctx, dout_own):
        from . import functional as F_
        g, key = ctx.g, ctx.key
        g_local = _halo_exchange(dout_own.contiguous(), g, ("smb", key) if key is not None else None)
        return F_.edge_gather_sum(g_local, g.i3, node_scale=g.dv_inv_local), None, None


class ShardedTwoHopFn(torch.autograd.Function):
    """Order-1 round trip over a partitioned hypergraph without the [E_local, d] intermediate:
    p_own [n_own, d] -> Dv^-1 * H H^T p for the own rows: halo_exchange, one two-hop pass over the local table
    (`ihg_two_hop_reduce`: replaces gather-sum + segmented sum), halo_reduce.  H H^T is symmetric, so backward
    is the same three steps with the scales swapped (Dv^-1 as the node scale, none on the rows)."""

    @staticmethod
    def _round_trip(x_own, g: "ShardedHyperGraph", key, node_scale, row_scale, tag: str):
        from . import _lib
        from . import functional as F_
        x_own = _lib.rows_f32(x_own)
        d = int(x_own.shape[1])
        x_local = _halo_exchange(x_own, g, (tag + "x", key) if key is not None else None)
        rkey = (tag + "s", key) if key is not None else None
        site = g.reduce_site(rkey, d)
        if site is not None:
            F_.two_hop_reduce_routed(g.plan_csr, g.two_hop_nbr, x_local, site.route, node_scale=node_scale)
            return _finish_routed(site, g, row_scale)
        s_local = F_.two_hop_reduce(g.plan_csr, g.two_hop_nbr, x_local, node_scale=node_scale,
                                    out=reduce_buffer(g, rkey, d, x_local))
        return _halo_reduce(s_local, g, row_scale, rkey)

    @staticmethod
    def forward(ctx, p_own, g: "ShardedHyperGraph", key=None):
        ctx.g, ctx.key = g, key
        return ShardedTwoHopFn._round_trip(p_own, g, key, None, g.dv_inv_own, "f")

    @staticmethod
    def backward(ctx, dout_own):
        g, key = ctx.g, ctx.key
        return ShardedTwoHopFn._round_trip(dout_own.contiguous(), g, key, g.dv_inv_local, None, "b"), None, None


def _halo_exchange(x_own: torch.Tensor, g: "ShardedHyperGraph", key=None) -> torch.Tensor:
    """[n_own, d] -> [n_local, d] = [own rows ; rows received from their owners] (no autograd).
    `key` names the call site: with peer memory the result lives in that site's persistent buffer."""
    from . import functional as F_
    from . import _lib
    d = int(x_own.shape[1])
    if g.p2p and key is not None:
        buf, hdl, _chunk, own, _off, n, chunk, off = g.peer_buffer(("x", key), d)
        x_own = _lib.rows_f32(x_own)
        # push: row send_rows[r] of my table -> my chunk in the reader's local table.  send_rows is
        # ordered by destination rank and holds nothing for myself, so the flat order matches `off`.
        src = (type(own))(*[x_own.data_ptr()] * n)
        x_local = buf[:g.n_local]
        _lib.call("ihg_halo_copy", src, chunk, off, n, _lib.ptr(g.send_rows_push), _lib.ld(x_own), d, d,
                  _lib.stream_ptr(), tag="halo_push", algo_bytes=g.S * (8 + 8 * d))
        if x_own.data_ptr() != x_local.data_ptr():  # producers may have written the own rows in place (table_head)
            F_.copy_rows_raw(x_own, x_local[:g.n_own])
        hdl.barrier(channel=0)                      # every rank's pushes have landed
        return x_local
    send = F_.gather_rows_raw(x_own, g.send_rows, 0)
    x_local = torch.empty((g.n_local, d), dtype=torch.float32, device=x_own.device)
    _all_to_all(x_local[g.n_own:], send, g.recv_counts, g.send_counts, g.group)
    F_.copy_rows_raw(x_own, x_local[:g.n_own])
    return x_local


class ShardedFeatureInteractFn(torch.autograd.Function):
    """Order 2/3 FeatureInteractor over a partitioned hypergraph, un-hoisted tensor-core forward:
    only the projected rows xp travel (d floats per boundary row); backward reduces the
    per-row gradients [dxp_hi | dP] of own + halo rows to their owners in ONE exchange and then
    applies the typed first-order Linear backward on the own rows."""

    @staticmethod
    def forward(ctx, xp_own, w_agg, bias, g: "ShardedHyperGraph", order: int, key=None):
        from . import _lib
        xp_own = _lib.rows_f32(xp_own)
        w_agg = _lib.rows_f32(w_agg)
        bias = bias.contiguous()
        dim, E = int(xp_own.shape[1]), g.EdgeCount
        ctx.key = key
        xp_local = _halo_exchange(xp_own, g, ("fi", key) if key is not None else None)
        ef = torch.empty((E, dim), dtype=torch.float32, device=xp_own.device)
        ws_bytes = _lib.lib().ihg_edge_interact_fwd_workspace_bytes(dim, order)
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=xp_own.device)
        _lib.call("ihg_feature_interact_fwd", _lib.ptr(xp_local), dim, _lib.ptr(w_agg), _lib.ld(w_agg),
                  _lib.ptr(bias), order, _lib.ptr(g.i3), E, _lib.ptr(ef), dim, dim, _lib.ptr(ws), ws_bytes,
                  _lib.stream_ptr(), tag="edge_interact_fwd", algo_bytes=E * (12 + 16 * dim))
        ctx.g, ctx.order = g, order
        ctx.save_for_backward(xp_own, xp_local, w_agg)
        return ef

    @staticmethod
    def backward(ctx, def_):
        from . import _lib
        from . import functional as F_
        from .layers import _split_first_order
        xp_own, xp_local, w_agg = ctx.saved_tensors
        g, order = ctx.g, ctx.order
        def_ = _lib.rows_f32(def_)
        dim, E = int(xp_own.shape[1]), g.EdgeCount
        nb = 4 if order == 3 else 3
        w_hi = w_agg[:, 3 * dim:]
        w_lo = _split_first_order(w_agg, dim).contiguous()
        slot_grad = torch.empty((E, 3, dim), dtype=torch.float32, device=def_.device)
        dw_hi = torch.empty((dim, nb * dim), dtype=torch.float32, device=def_.device)
        ws_bytes = _lib.lib().ihg_edge_interact_bwd_workspace_bytes(dim, order)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=def_.device)
        _lib.call("ihg_edge_interact_bwd", _lib.ptr(xp_local), dim, _lib.ptr(def_), _lib.ld(def_),
                  _lib.ptr(w_hi), _lib.ld(w_hi), order, _lib.ptr(g.i3), E, _lib.ptr(slot_grad),
                  _lib.ptr(dw_hi), dim, _lib.ptr(ws), ws_bytes, _lib.stream_ptr(),
                  tag="edge_interact_bwd", algo_bytes=E * (12 + 28 * dim))
        # per-row gradients of the local rows, side by side: [ product-rule part | dP ]
        rkey = ("fib", ctx.key) if ctx.key is not None else None
        site = g.reduce_site(rkey, 2 * dim)
        if site is not None:
            F_.segment_reduce_routed(g.plan_csr, slot_grad, dim, site.route, src_row_mul=3, row_slot=g.row_slot)
            F_.segment_reduce_routed(g.plan_csr, def_, dim, site.route.shifted(4 * dim))
            own = _finish_routed(site, g, None)                            # [n_own, 2 dim]
        else:
            both = reduce_buffer(g, rkey, 2 * dim, def_)
            F_.segment_reduce(g.plan_csr, slot_grad, dim, src_row_mul=3, row_slot=g.row_slot, out=both[:, :dim])
            F_.segment_reduce(g.plan_csr, def_, dim, out=both[:, dim:])
            own = _halo_reduce(both, g, None, rkey)                        # [n_own, 2 dim]
        dxp_hi, dp = own[:, :dim], own[:, dim:]
        dxp = F_.node_linear(dp, w_lo, transpose_w=True, addend=dxp_hi, bounds=g.own_bounds)
        dw_lo, db_lo = F_.node_linear_wgrad(dp, xp_own, 3, g.own_bounds, True)
        dw = torch.cat([dw_lo[0], dw_lo[1], dw_lo[2], dw_hi], 1)
        return dxp, dw, db_lo[0], None, None, None


def _halo_reduce(s_local: torch.Tensor, g: ShardedHyperGraph, row_scale: Optional[torch.Tensor], key=None) -> torch.Tensor:
    """own rows <- row_scale * (own partial + partials received from the ranks holding them as halo
    rows, ascending source rank).  The halo partials are the contiguous tail of s_local.
    With peer memory (`key` names the call site and s_local must be that site's peer buffer, see
    `reduce_buffer`) the owner pulls the partials straight out of the holders' buffers."""
    from . import functional as F_
    from . import _lib
    d = int(s_local.shape[1])
    if g.p2p and key is not None:
        buf, hdl, chunk, own, off, n = g.peer_buffer(("s", key), d)[:6]
        assert s_local.data_ptr() == buf.data_ptr(), "halo_reduce: s_local must be the call site's peer buffer"
        out = torch.empty((g.n_own, d), dtype=torch.float32, device=s_local.device)
        hdl.barrier(channel=0)                      # every rank's partial sums are complete
        # one kernel: own partial + the holders' partials read straight over NVLink, ascending rank, scaled
        _lib.call("ihg_halo_reduce", _lib.ptr(s_local), d, _lib.ptr(g.reduce_rowptr), _lib.ptr(g.reduce_entries),
                  chunk, n, d, _lib.ptr(row_scale), _lib.ptr(out), d, g.n_own, d, _lib.stream_ptr(),
                  tag="halo_reduce", algo_bytes=g.S * 4 * d + g.n_own * 8 * d)
        if g.share_buffers:
            hdl.barrier(channel=0)                  # holders may overwrite the (shared) buffer again; a per-site buffer
        return out                                  # is next written a whole step -- many barriers -- later
    recv = torch.empty((g.S, d), dtype=torch.float32, device=s_local.device)
    _all_to_all(recv, s_local[g.n_own:], g.send_counts, g.recv_counts, g.group)
    return F_.segment_reduce(g.reduce_csr, recv, d, row_scale=row_scale, init=s_local[:g.n_own])


def _finish_routed(site: "_ReduceSite", g: ShardedHyperGraph, row_scale: Optional[torch.Tensor]) -> torch.Tensor:
    """Second half of a routed reduce: every rank's reduction kernel has written its partial sums of my rows
    into `site.recv` (posted NVLink stores); after the barrier the owner adds its own partial and the received
    ones in ascending source rank (`reduce_csr` over the flat receive layout), applies `row_scale`."""
    from . import functional as F_
    d = int(site.own.shape[1])
    site.hdl.barrier(channel=0)                     # all ranks' partial sums have landed
    if g.S == 0:
        return site.own * row_scale.view(-1, 1) if row_scale is not None else site.own.clone()
    # a per-site receive buffer is next written a whole step -- many barriers -- later: no second barrier
    return F_.segment_reduce(g.reduce_csr, site.recv, d, row_scale=row_scale, init=site.own)


def halo_exchange(x_own, g, key=None):
    return HaloExchangeFn.apply(x_own, g, key)


def reduce_buffer(g: ShardedHyperGraph, key, cols: int, like: torch.Tensor) -> torch.Tensor:
    """[n_local, cols] buffer for partial sums that will go through `_halo_reduce(.., key=key)`:
    the call site's peer buffer when peer memory is on, a fresh tensor otherwise."""
    if g.p2p and key is not None:
        return g.peer_buffer(("s", key), cols)[0][:g.n_local]
    return torch.empty((g.n_local, cols), dtype=torch.float32, device=like.device)


class _LocalGraphView:
    """What the edge / segment kernels' autograd Functions read from a graph object."""

    def __init__(self, g: ShardedHyperGraph):
        self.i3, self.plan, self.EdgeCount = g.i3, g.plan_csr, g.EdgeCount
        self.type_bounds = (1 << 62, 1 << 62)          # unused: row_slot gives the slot of a local row
        self.row_slot = g.row_slot
        self.dv_inv = None


class ShardedIHGNNLayer(torch.nn.Module):
    """IHGNNLayer (Models/GnnLayers.py:156-236) over a partitioned hypergraph.  Holds the same
    parameters (`feature_interactor.aggregation`, `feature_transform`) as the single-GPU layer;
    input / output are the rank's own rows [n_own, d].  Dense weight gradients must be
    all-reduced by the caller (`allreduce_dense_grads`)."""

    _count = 0

    def __init__(self, graph: ShardedHyperGraph, dim: int, feature_interaction_order: int):
        super().__init__()
        from .layers import FeatureInteractor

        class _DS:           # the attribute FeatureInteractor reads
            pass
        ds = _DS()
        ds.graph = _LocalGraphView(graph)
        ds.graph.type_bounds = graph.own_bounds        # FeatureInteractor's typed Linear runs on own rows
        self.g = graph
        self.view = _LocalGraphView(graph)
        self.order = feature_interaction_order
        ShardedIHGNNLayer._count += 1
        self.uid = ShardedIHGNNLayer._count           # names this layer's persistent peer-memory buffers
        self.feature_interactor = FeatureInteractor(ds, feature_interaction_order, dim, dim)
        self.feature_transform = torch.nn.Linear(dim, dim)

    def forward(self, x_own: torch.Tensor) -> torch.Tensor:
        from . import functional as F_
        from .layers import _EdgeGatherSumFn, _EdgeInteractFn, _split_first_order
        g, fi = self.g, self.feature_interactor
        wt, bt = self.feature_transform.weight, self.feature_transform.bias
        d = fi.node_feature_dimension
        w_lo = _split_first_order(fi.aggregation.weight, d)
        zeros = torch.zeros_like(bt)
        if self.order == 1:
            w_f = torch.matmul(w_lo, wt)
            b_f = torch.matmul(w_lo, bt) + torch.stack([fi.aggregation.bias, zeros, zeros])
            if F_.two_hop_enabled(g.n_local, d):
                p_own = F_.typed_linear(x_own, w_f, b_f, g.own_bounds, out=g.table_head(("fx", self.uid), d))
                return ShardedTwoHopFn.apply(p_own, g, self.uid)
            p_own = F_.typed_linear(x_own, w_f, b_f, g.own_bounds, out=g.table_head((self.uid, "p"), d))
            p = halo_exchange(p_own, g, (self.uid, "p"))
            ef = _EdgeGatherSumFn.apply(p, self.view, None, 1.0, None)
        else:
            from . import _lib
            tc = bool(_lib.lib().ihg_feature_interact_supported(d))
            xp_own = F_.typed_linear(x_own, wt.unsqueeze(0), bt.unsqueeze(0), None,
                                     out=g.table_head(("fi", self.uid), d) if tc else None)
            if tc:
                ef = ShardedFeatureInteractFn.apply(xp_own, fi.aggregation.weight, fi.aggregation.bias, g, self.order, self.uid)
            else:
                b_lo = torch.stack([fi.aggregation.bias, zeros, zeros])
                p_own = F_.typed_linear(xp_own, w_lo, b_lo, g.own_bounds)
                both = halo_exchange(torch.cat([xp_own, p_own], 1), g, (self.uid, "xpp"))     # one exchange for both row sets
                xp, p = both[:, :d], both[:, d:]
                ef = _EdgeInteractFn.apply(xp, p, fi.aggregation.weight[:, 3 * d:], self.view, self.order)
        return ShardedScatterMeanFn.apply(ef, g, self.uid)


def allreduce_dense_grads(module: torch.nn.Module, group=None) -> None:
    """Sum the (small, dense) weight gradients over ranks: every rank saw only its hyperedges."""
    import torch.distributed as dist
    grads = [p.grad for p in module.parameters() if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n


# ----------------------------------------------------------------------------------------
# sharded model: row-sharded embedding tables + sharded conv stack + replicated scorer
# ----------------------------------------------------------------------------------------
class _FetchRowsFn(torch.autograd.Function):
    """Rows of a row-sharded matrix by GLOBAL id, available on every rank: out[j] = F[global_ids[j]].
    Fixed-shape and sync-free: every rank gathers ONE row per requested id -- its own row when it owns
    the id (mask 1), an arbitrary own row otherwise (mask 0) -- zeroes the rows it does not own and one
    all-reduce(sum) completes the matrix (the batch head needs only B x 3 rows, SURVEY 8e "tiny").
    Backward: every rank already holds the full gradient; it scatter-adds the masked rows (the ones it
    does not own contribute exact zeros; duplicates summed in ascending j: deterministic)."""

    @staticmethod
    def forward(ctx, f_own, local_rows, mask, group):
        import torch.distributed as dist
        from . import functional as F_
        out = F_.gather_rows_raw(f_own, local_rows, 0)
        out.mul_(mask.view(-1, 1))
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
        ctx.save_for_backward(local_rows, mask)
        ctx.shape = tuple(f_own.shape)
        return out

    @staticmethod
    def backward(ctx, dout):
        from . import _lib
        local_rows, mask = ctx.saved_tensors
        d = int(dout.shape[1])
        g = (dout * mask.view(-1, 1)).contiguous()
        df = torch.zeros(ctx.shape, dtype=torch.float32, device=dout.device)
        _lib.call("ihg_scatter_add_rows", _lib.ptr(g), d, _lib.ptr(local_rows), 0, int(local_rows.numel()),
                  _lib.ptr(df), d, d, _lib.stream_ptr())
        return df, None, None, None


class ShardedRawGnn(torch.nn.Module):
    """RawGnn (Models/RawGnn.py:14-144) over a partitioned hypergraph: user / item embedding rows
    are sharded like the nodes, the vocabulary table, conv weights and item biases are replicated
    (their gradients are all-reduced by `sync_grads`).  `forward` takes the GLOBAL batch indices
    (identical on every rank) and returns the scores of the whole batch on every rank."""

    def __init__(self, graph: ShardedHyperGraph, bag_words: np.ndarray, bag_offsets: np.ndarray,
                 vocab_size: int, embedding_size: int, layer_count: int, feature_interaction_order: int,
                 lambda_muq: float = 0.5):
        super().__init__()
        from .graph import CsrPlan, csr_from_keys
        from .layers import HemPredictionLayer
        p, dev = graph.plan, graph.device
        self.g, self.lambda_muq = graph, lambda_muq
        d = embedding_size
        self.embedding_user = torch.nn.Parameter(torch.empty(p.Uo, d))
        self.embedding_item = torch.nn.Parameter(torch.empty(p.Io, d))
        self.embedding_bag_vocabulary = torch.nn.Parameter(torch.empty(vocab_size + 1, d))
        for w, rows in ((self.embedding_user, p.U + 1), (self.embedding_item, p.I + 1), (self.embedding_bag_vocabulary, vocab_size + 1)):
            bound = (6.0 / (rows + d)) ** 0.5               # xavier_uniform_ of the full table (fan_out = rows)
            torch.nn.init.uniform_(w, -bound, bound)
        self.gnns = torch.nn.ModuleList()
        for k in range(layer_count):
            order = feature_interaction_order if (k == 0 or feature_interaction_order == 1) else 1
            self.gnns.append(ShardedIHGNNLayer(graph, d, order))
        self.prediction_layer = HemPredictionLayer(d * (1 + layer_count), lambda_muq, p.I)
        self.to(dev)
        self.sync_replicated_parameters()
        # bag CSR of the own queries (+ its transpose) for EmbeddingBag(mean)
        words = np.asarray(bag_words, dtype=np.int64)
        ptr = np.concatenate([np.asarray(bag_offsets, dtype=np.int64), [words.shape[0]]])
        q0, q1 = int(p.qb[p.rank]), int(p.qb[p.rank + 1])
        own_words = torch.as_tensor(words[ptr[q0]:ptr[q1]], dtype=torch.int32, device=dev)
        own_ptr = torch.as_tensor(ptr[q0:q1 + 1] - ptr[q0], dtype=torch.int32, device=dev)
        lens = (own_ptr[1:] - own_ptr[:-1]).to(torch.float32)
        self.bag_inv_len = torch.where(lens > 0, 1.0 / lens.clamp(min=1.0), torch.zeros_like(lens))
        self.bag_plan = CsrPlan(own_ptr, own_words)
        bag_of = torch.repeat_interleave(torch.arange(p.Qo, device=dev, dtype=torch.int32), (own_ptr[1:] - own_ptr[:-1]).to(torch.int64))
        wptr, _perm, bags = csr_from_keys(own_words, vocab_size + 1, values=bag_of)
        self.word_plan = CsrPlan(wptr, bags)

    def input_features(self) -> torch.Tensor:
        from .layers import _BagMeanFn

        class _T:
            pass
        t = _T()
        t.bag_plan, t.word_plan, t.bag_inv_len = self.bag_plan, self.word_plan, self.bag_inv_len
        q = _BagMeanFn.apply(self.embedding_bag_vocabulary, t)
        return torch.cat([self.embedding_user, q, self.embedding_item])

    def output_features(self) -> torch.Tensor:
        h = self.input_features()
        outs = [h]
        for gnn in self.gnns:
            h = gnn(h)
            outs.append(h)
        return torch.cat(outs, 1)

    def forward(self, users: torch.Tensor, queries: torch.Tensor, items: torch.Tensor) -> torch.Tensor:
        p = self.g.plan
        f_own = self.output_features()
        B = int(users.numel())
        local_rows, mine = p.batch_rows(users, queries, items)
        rows = _FetchRowsFn.apply(f_own, local_rows, mine.to(torch.float32), self.g.group)
        return self.prediction_layer(rows[:B], rows[B:2 * B], rows[2 * B:], items)

    @torch.no_grad()
    def sync_replicated_parameters(self) -> None:
        """Rank 0's values for every REPLICATED parameter (conv weights, vocabulary table, item
        bias): the row-sharded tables differ in size per rank, so the ranks' RNG streams diverge
        during construction and the replicas would start out different."""
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return
        group = self.g.group
        src = dist.get_global_rank(group, 0) if group is not None else 0
        for q in list(self.gnns.parameters()) + [self.embedding_bag_vocabulary, self.prediction_layer.items_bias]:
            dist.broadcast(q.data, src=src, group=group)

    @torch.no_grad()
    def gather_features(self) -> torch.Tensor:
        """Replicate the output features on every rank in GLOBAL node order [U+Q+I, d(1+L)]
        (users, queries, items; Helpers/Graph.py:110-111): one padded all-gather of the own rows,
        then three row-range copies per source rank.  Inference only (BASELINE.json configs[4]:
        item features replicated, queries split across the GPUs, no per-query communication)."""
        import torch.distributed as dist
        from . import functional as F_
        p = self.g.plan
        f_own = self.output_features().contiguous()
        D = int(f_own.shape[1])
        n_max = max(int(p.ub[r + 1] - p.ub[r]) + int(p.qb[r + 1] - p.qb[r]) + int(p.ib[r + 1] - p.ib[r])
                    for r in range(p.world))
        send = torch.zeros((n_max, D), dtype=torch.float32, device=f_own.device)
        F_.copy_rows_raw(f_own, send[:p.n_own])
        recv = torch.empty((p.world, n_max, D), dtype=torch.float32, device=f_own.device)
        dist.all_gather_into_tensor(recv.view(-1, D), send, group=self.g.group)
        feat = torch.empty((p.U + p.Q + p.I, D), dtype=torch.float32, device=f_own.device)
        for r in range(p.world):
            uo, qo, io = int(p.ub[r + 1] - p.ub[r]), int(p.qb[r + 1] - p.qb[r]), int(p.ib[r + 1] - p.ib[r])
            for n, src0, dst0 in ((uo, 0, int(p.ub[r])), (qo, uo, p.U + int(p.qb[r])),
                                  (io, uo + qo, p.U + p.Q + int(p.ib[r]))):
                if n:
                    F_.copy_rows_raw(recv[r, src0:src0 + n], feat[dst0:dst0 + n])
        return feat

    @torch.no_grad()
    def rank(self, users: torch.Tensor, queries: torch.Tensor, candidates: Optional[torch.Tensor] = None,
             k: int = 10, features: Optional[torch.Tensor] = None):
        """Rank THIS rank's share of the searches: rows [rank::world] of (users, queries, candidates)
        against the replicated features (`gather_features`, pass it in to reuse it across batches).
        Returns (row indices of the share, item ids [n, k], scores [n, k]); no collective per query."""
        from . import functional as F_
        p = self.g.plan
        feat = self.gather_features() if features is None else features
        share = torch.arange(p.rank, int(queries.numel()), p.world, device=feat.device)
        cand = candidates[share] if candidates is not None else None
        pl = self.prediction_layer
        from .settings import Gs
        ids, vals = F_.rank_topk(feat, users[share], queries[share], pl.items_bias, pl.lambda_muq,
                                 query_row0=p.U, item_row0=p.U + p.Q, item_count=p.I, candidates=cand, k=k,
                                 cosine=bool(Gs.Prediction.use_cosine_similarity))
        return share, ids, vals

    def sync_grads(self) -> None:
        """All-reduce the gradients of the replicated parameters whose per-rank gradients are
        partial sums (conv weights, vocabulary table).  items_bias sees identical gradients on
        every rank (the scores are computed redundantly), so it needs no reduction."""
        import torch.distributed as dist
        grads = [q.grad for q in list(self.gnns.parameters()) + [self.embedding_bag_vocabulary] if q.grad is not None]
        if not grads:
            return
        flat = torch.cat([g.reshape(-1) for g in grads])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.g.group)
        off = 0
        for g in grads:
            n = g.numel()
            g.copy_(flat[off:off + n].view_as(g))
            off += n
