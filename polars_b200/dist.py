"""Multi-GPU plumbing for the partitioned plan: one process per GPU,
`torch.distributed` for the exchange.  The data path has exactly one exchange step — an all-to-all-v of
hash-partitioned rows (partial aggregates for group_by, (key,row) pairs for join).

The reference has no distributed path; the partition function is its in-memory one
(hash_to_partition(dirty_hash(key), P), crates/polars-utils/src/hashing.rs:62-69,132-142), which
the device kernels restate bit-exactly, so per-destination counts are checkable integers.

This module is host logic only (split sizes, buffers, collective calls) and is backend-agnostic:
NCCL on GPUs (`all_to_all_single`), gloo on CPU for the world_size-2 tests (isend/irecv pairs,
because gloo has no all_to_all).
"""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist


def exchange_counts(send_counts: np.ndarray, device) -> np.ndarray:
    """send_counts[p] = rows this rank sends to rank p  ->  recv_counts[p] = rows rank p sends here."""
    world = dist.get_world_size()
    assert send_counts.shape == (world,)
    sc = torch.tensor(send_counts, dtype=torch.int64, device=device)
    if dist.get_backend() == "nccl":
        rc = torch.empty_like(sc)
        dist.all_to_all_single(rc, sc)
    else:
        gathered = [torch.empty_like(sc) for _ in range(world)]
        dist.all_gather(gathered, sc)
        rc = torch.stack([g[dist.get_rank()] for g in gathered])
    return rc.cpu().numpy()


def all_to_all_rows(send: torch.Tensor, send_counts: np.ndarray, row_words: int):
    """send: int64 tensor of sum(send_counts)*row_words words, destination-major (partition p's rows
    are contiguous).  Returns (recv tensor, recv_counts)."""
    world, rank = dist.get_world_size(), dist.get_rank()
    recv_counts = exchange_counts(np.asarray(send_counts, dtype=np.int64), send.device)
    recv = torch.empty(int(recv_counts.sum()) * row_words, dtype=torch.int64, device=send.device)
    in_split = (np.asarray(send_counts, dtype=np.int64) * row_words).tolist()
    out_split = (recv_counts * row_words).tolist()
    if dist.get_backend() == "nccl":
        dist.all_to_all_single(recv, send, output_split_sizes=out_split, input_split_sizes=in_split)
    else:
        so = np.concatenate([[0], np.cumsum(in_split)])
        ro = np.concatenate([[0], np.cumsum(out_split)])
        recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
        reqs = []
        for p in range(world):
            if p == rank:
                continue
            if out_split[p]:
                reqs.append(dist.irecv(recv[ro[p]:ro[p + 1]], src=p))
            if in_split[p]:
                reqs.append(dist.isend(send[so[p]:so[p + 1]].contiguous(), dst=p))
        for r in reqs:
            r.wait()
    return recv, recv_counts


class CudaWords:
    """Zero-copy view of a raw device pointer as int64 words (CUDA array interface v2) for torch.as_tensor."""

    def __init__(self, ptr: int, n_words: int):
        self.__cuda_array_interface__ = {"shape": (int(n_words),), "typestr": "<i8", "data": (int(ptr), False), "version": 2}


def partitioned_group_by(plb, key_col, value_cols, spec, location=None, nullable=None):
    """One rank's share of the partitioned group_by: local pre-aggregation (K5) -> hash-partitioned
    export of the partial aggregates (K6) -> one all-to-all -> merge (K5 merge) -> finish.
    Output stays partitioned: this rank owns the groups with hash_to_partition(key) == rank."""
    world = dist.get_world_size()
    g = plb.GroupBy(plb.NP_OF[key_col.dtype], spec, nullable=nullable)
    g.consume(key_col, value_cols, row_base=0)
    ptr, rw, offs = g.export_partials(world)
    try:
        n_send = int(offs[-1])
        send = torch.as_tensor(CudaWords(ptr, n_send * rw), device="cuda") if n_send else torch.empty(0, dtype=torch.int64, device="cuda")
        recv, rc = all_to_all_rows(send, np.diff(offs), rw)
        torch.cuda.synchronize()
        f = plb.GroupBy(plb.NP_OF[key_col.dtype], spec, expected_groups=max(int(rc.sum()), 1), nullable=nullable)
        f.merge_partials(recv.data_ptr(), int(rc.sum()))
        return f.finish(False, location=plb.DEVICE if location is None else location)
    finally:
        plb.dev_free(ptr)


WINDOW_HEADER_BYTES = 1024      # BL_WINDOW_HEADER_BYTES: per source rank (rows sent, epoch flag)


class PeerExchange:
    """Peer windows for the fused partition + exchange (K6 stores straight into the destination GPU's
    memory over NVLink; the row counts and a ready flag travel through the window headers, so the data
    path has no collective and no host round trip at all).
    Two window halves alternate between steps: a rank that already runs step k+1 writes into the other
    half than the one a slow rank is still merging from step k; it can only reach step k+2 (same half
    again) after its own merge of step k+1, which waits for the slow rank's step-k+1 flags — and those are
    published by the slow rank's export kernel, stream-ordered after its merge of step k.  So two halves
    are enough and no barrier is needed."""

    def __init__(self, plb, rows_per_src: int, row_words: int):
        self.plb, self.rows_per_src, self.row_words = plb, int(rows_per_src), int(row_words)
        self.world, self.rank = dist.get_world_size(), dist.get_rank()
        self.half_bytes = WINDOW_HEADER_BYTES + self.world * self.rows_per_src * self.row_words * 8
        self.half_bytes = (self.half_bytes + 255) // 256 * 256
        self.epoch = 0
        # every collective below is reached by every rank even if a local step fails (no rank may hang)
        try:
            self.win, err = plb.Window(2 * self.half_bytes), None
        except Exception as e:
            self.win, err = None, str(e)
        handles = [None] * self.world
        dist.all_gather_object(handles, None if self.win is None else self.win.ipc_handle)
        self.peers = []
        if all(h is not None for h in handles):
            try:
                self.peers = [self.win.ptr if r == self.rank else plb.Window.open(handles[r]) for r in range(self.world)]
            except Exception as e:
                err = str(e)
        else:
            err = err or "a peer could not create its window"
        flags = [None] * self.world
        dist.all_gather_object(flags, err)
        if any(f is not None for f in flags):
            if self.win is not None:
                self.win.destroy()
            raise RuntimeError("peer windows unavailable: " + "; ".join(str(f) for f in flags if f))
        self.half = 0

    def next_step(self) -> int:
        """Flips the window half and returns the step's epoch (strictly increasing; flags are never reset)."""
        self.half ^= 1
        self.epoch += 1
        return self.epoch

    def peer_halves(self):
        return [p + self.half * self.half_bytes for p in self.peers]

    def own_half(self) -> int:
        return self.win.ptr + self.half * self.half_bytes

    def close(self):
        dist.barrier()
        for r, p in enumerate(self.peers):
            if r != self.rank:
                self.plb.Window.close(p)
        self.win.destroy()


def partitioned_group_by_p2p(plb, ex: PeerExchange, key_col, value_cols, spec, location=None, nullable=None, expected_groups: int = 0):
    """Same plan as partitioned_group_by, but the partial aggregates reach their owner by P2P stores from inside
    the partition kernel and the owner's merge kernel picks the row counts up from its window header on the
    device: local K5 -> export (stores + publish) -> merge (wait + merge) -> finish are queued back to back by ONE
    library call (bl_groupby_agg_partitioned); the only host synchronisations of the step are the cardinality sample of
    the local K5 and the final read-back of the group count.  Errors of the deferred checks (local table overflow, peer
    region overflow, peer timeout) are raised at the end of the step.  spec: [(kind, dtype | None)], value_cols aligned."""
    epoch = ex.next_step()
    aggs = [(kind, v) for (kind, _), v in zip(spec, value_cols)]
    return plb.group_by_agg_partitioned(key_col, aggs, ex.rank, ex.peer_halves(), ex.own_half(), ex.rows_per_src, epoch, expected_groups,
                                        location=plb.DEVICE if location is None else location)


class _CudaArr:
    def __init__(self, ptr: int, n: int, typestr: str):
        self.__cuda_array_interface__ = {"shape": (int(n),), "typestr": typestr, "data": (int(ptr), False), "version": 2}


def _as_torch(out_col, n: int, typestr: str, dtype):
    return torch.as_tensor(_CudaArr(out_col.values_ptr, n, typestr), device="cuda") if n else torch.empty(0, dtype=dtype, device="cuda")


def all_to_all_1d(send: torch.Tensor, send_counts: np.ndarray):
    """all-to-all-v of a 1-D tensor whose destination-p slice has send_counts[p] elements."""
    recv_counts = exchange_counts(np.asarray(send_counts, dtype=np.int64), send.device)
    recv = torch.empty(int(recv_counts.sum()), dtype=send.dtype, device=send.device)
    dist.all_to_all_single(recv, send, output_split_sizes=recv_counts.tolist(), input_split_sizes=np.asarray(send_counts).tolist())
    return recv, recv_counts


def partitioned_hash_join(plb, left_key, right_key, left_base: int, right_base: int, how: str = "inner"):
    """Partitioned inner/left hash join on P GPUs: both sides are hash-partitioned on the key
    (K6, the reference's partition function), exchanged with ONE count exchange + one all-to-all-v per column and
    relation, joined locally (K7/K8) and mapped back to GLOBAL row ids (K4).  left_key/right_key: device Columns of this
    rank's rows; *_base: global index of this rank's first row.  Returns (left_global_idx, right_global_idx) device
    columns; the output is partitioned by key."""
    world = dist.get_world_size()
    sides = []
    for key, base in ((left_key, left_base), (right_key, right_base)):
        n = key.length
        gid = torch.arange(base, base + n, dtype=torch.int32, device="cuda") if base + n < 2**31 else (torch.arange(n, dtype=torch.int64, device="cuda") + base).to(torch.int32)
        torch.cuda.current_stream().synchronize()
        gcol = plb.Column(gid.data_ptr(), dtype=np.uint32, length=n, location=plb.DEVICE)
        kp, [gp], offs = plb.hash_partition(key, [gcol], world, location=plb.DEVICE)
        counts = np.diff(offs)
        kt = _as_torch(kp, n, "<i8", torch.int64)
        gt = _as_torch(gp, n, "<i4", torch.int32)
        (rk, rg), _ = exchange_columns([kt, gt], counts)
        sides.append((rk, rg, kp, gp))
    torch.cuda.synchronize()
    (lk, lg, *_), (rk, rg, *_) = sides
    lcol = plb.Column(lk.data_ptr(), dtype=np.int64, length=lk.numel(), location=plb.DEVICE)
    rcol = plb.Column(rk.data_ptr(), dtype=np.int64, length=rk.numel(), location=plb.DEVICE)
    li, ri = plb.hash_join(lcol, rcol, how, False, "none", location=plb.DEVICE)
    lgc = plb.Column(lg.data_ptr(), dtype=np.uint32, length=lg.numel(), location=plb.DEVICE)
    rgc = plb.Column(rg.data_ptr(), dtype=np.uint32, length=rg.numel(), location=plb.DEVICE)
    [gl] = plb.gather([lgc], li.view(), check_bounds=False, location=plb.DEVICE)
    [gr] = plb.gather([rgc], ri.view(), check_bounds=False, location=plb.DEVICE)
    return gl, gr


def broadcast_hash_join(plb, left_key, right_key, left_base: int, how: str = "inner"):
    """Small-build-side alternative: every rank gathers the WHOLE build (right) key column with
    one all-gather and joins its own probe (left) rows against it — no probe-side exchange at all (C3: 80 MB of
    build keys against 175 MB of probe rows per GPU).  Rank r's build rows must be rows [r*n_r, (r+1)*n_r) of the
    global build relation (equal n_r on every rank), so the gathered order IS the global row id and the tuples
    come back in global ids directly.  Returns (left_global_idx, right_global_idx) device columns; the output
    stays distributed by probe row (rank r holds the matches of its own rows, in the reference's probe order)."""
    world = dist.get_world_size()
    n_r = right_key.length
    mine = _as_torch(right_key, n_r, "<i8", torch.int64) if hasattr(right_key, "values_ptr") else \
        torch.as_tensor(_CudaArr(right_key._vptr, n_r, "<i8"), device="cuda")
    full = torch.empty(n_r * world, dtype=torch.int64, device="cuda")
    dist.all_gather_into_tensor(full, mine)
    torch.cuda.synchronize()
    rcol = plb.Column(full.data_ptr(), dtype=np.int64, length=full.numel(), location=plb.DEVICE)
    li, ri = plb.hash_join(left_key, rcol, how, False, "none", location=plb.DEVICE)
    # local probe idx -> global: add this rank's base (K1)
    base = plb.Column(np.array([left_base], np.uint32))
    gl = plb.elementwise("add", li.view(), base, location=plb.DEVICE)
    return gl, ri


def choose_join_plan(probe_rows: int, build_rows: int, world: int) -> str:
    """The exchange-volume rule a planner would apply (per GPU, bytes over NVLink): the partitioned plan ships
    (world-1)/world of BOTH relations as (key, row id) records (12 B per row); the broadcast plan receives the other ranks'
    build keys only (8 B per row) and leaves the probe rows where they are.  C3 (1e8 x 1e7 per GPU): 1.1 GB vs 0.56 GB at 8
    GPUs — broadcast; two relations of similar size: partitioned."""
    part = (world - 1) / world * (probe_rows + build_rows) * 12
    bcast = (world - 1) * build_rows * 8
    return "broadcast" if bcast <= part else "partitioned"


def exchange_columns(cols, send_counts: np.ndarray):
    """all-to-all-v of several 1-D tensors that share one destination-major row layout (partition p's rows are
    contiguous in every column): ONE count exchange, then one all-to-all per column.  Backend-agnostic (NCCL on
    GPUs, isend/irecv pairs under gloo).  Returns ([received columns], recv_counts)."""
    send_counts = np.asarray(send_counts, dtype=np.int64)
    world, rank = dist.get_world_size(), dist.get_rank()
    recv_counts = exchange_counts(send_counts, cols[0].device if cols else "cpu")
    so, ro = np.concatenate([[0], np.cumsum(send_counts)]), np.concatenate([[0], np.cumsum(recv_counts)])
    outs = []
    for send in cols:
        assert send.numel() == int(send_counts.sum())
        recv = torch.empty(int(recv_counts.sum()), dtype=send.dtype, device=send.device)
        if dist.get_backend() == "nccl":
            dist.all_to_all_single(recv, send, output_split_sizes=recv_counts.tolist(), input_split_sizes=send_counts.tolist())
        else:
            recv[ro[rank]:ro[rank + 1]] = send[so[rank]:so[rank + 1]]
            reqs = []
            for p in range(world):
                if p == rank:
                    continue
                if recv_counts[p]:
                    reqs.append(dist.irecv(recv[ro[p]:ro[p + 1]], src=p))
                if send_counts[p]:
                    reqs.append(dist.isend(send[so[p]:so[p + 1]].contiguous(), dst=p))
            for r in reqs:
                r.wait()
        outs.append(recv)
    return outs, recv_counts


def partitioned_group_by_rows(plb, key_col, value_cols, aggs, location=None):
    """High-cardinality plan: when local pre-aggregation would not shrink the data (groups ~
    rows), the RAW rows are hash-partitioned on the key (K6: key + value columns scattered together), exchanged with
    one all-to-all-v per column, and aggregated once on the owning rank (K5).  `aggs`: [(kind, index into value_cols
    | None)].  8-byte key and value columns without nulls.  Output stays partitioned by key."""
    world = dist.get_world_size()
    n = key_col.length
    kp, vps, offs = plb.hash_partition(key_col, value_cols, world, location=plb.DEVICE)
    counts = np.diff(offs)
    key_t = _as_torch(kp, n, "<i8", torch.int64)
    val_t = [_as_torch(v, n, "<i8", torch.int64) for v in vps]
    recv, rc = exchange_columns([key_t] + val_t, counts)
    torch.cuda.synchronize()
    m = int(rc.sum())
    kcol = plb.Column(recv[0].data_ptr(), dtype=plb.NP_OF[key_col.dtype], length=m, location=plb.DEVICE)
    vcols = [plb.Column(r.data_ptr(), dtype=plb.NP_OF[v.dtype], length=m, location=plb.DEVICE) for r, v in zip(recv[1:], value_cols)]
    return plb.group_by_agg(kcol, [(kind, None if i is None else vcols[i]) for kind, i in aggs], False, location=plb.DEVICE if location is None else location)
