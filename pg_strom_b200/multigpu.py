"""One process per GPU: how the table's chunks are dealt to the ranks and how
the per-GPU partial states come together.

The reference deals messages round-robin to its OpenCL devices
(opencl_serv.c:100-106) and lets PostgreSQL's final Agg merge the partial
rows (gpupreagg.c:2169-2186).  Here chunk i belongs to rank i mod G, every
rank keeps one persistent state for the whole scan, and before the final Agg
the states are merged into the root over NVLink with NCCL
(pgs_preagg_merge_nccl in the C ABI).  torch.distributed is plumbing only: it
carries the 128-byte NCCL id to the ranks and the barrier / max-over-ranks of
the measurements; the data path never goes through it.
"""
import ctypes as C


def deal_chunks(nchunks, rank, world):
    """Chunk indices of `rank`: round-robin, chunk i -> rank i mod world."""
    if not (0 <= rank < world):
        raise ValueError("rank %d outside world of %d" % (rank, world))
    return list(range(rank, nchunks, world))


def shard_rows(nrows, chunk_rows, rank, world):
    """[(row0, n)] of the chunks of a table of `nrows` rows that `rank` scans."""
    nchunks = (nrows + chunk_rows - 1) // chunk_rows
    return [(i * chunk_rows, min(chunk_rows, nrows - i * chunk_rows))
            for i in deal_chunks(nchunks, rank, world)]


def broadcast_bytes(dist, payload, nbytes, src=0, device=None):
    """Every rank gets the `nbytes` bytes rank `src` passes as `payload`
    (works on any torch.distributed backend; used for the NCCL unique id)."""
    import torch
    if dist.get_rank() == src:
        raw = bytes(payload)[:nbytes].ljust(nbytes, b"\0")
        t = torch.tensor(list(raw), dtype=torch.uint8)
    else:
        t = torch.zeros(nbytes, dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    dist.broadcast(t, src)
    return bytes(t.cpu().numpy().tobytes())


def nccl_communicator(lib, dist, local_device_index, rank, world, device=None):
    """ncclComm_t (as c_void_p) shared by the ranks of `dist`'s default group."""
    from . import _capi
    uid = C.create_string_buffer(128)
    if rank == 0:
        _capi.check(lib.pgs_nccl_get_unique_id(uid))
    raw = broadcast_bytes(dist, uid.raw, 128, 0, device)
    uid = C.create_string_buffer(raw, 128)
    comm = C.c_void_p()
    _capi.check(lib.pgs_nccl_comm_init_rank(local_device_index, world, uid, rank, C.byref(comm)))
    return comm


def max_over_ranks(dist, value, device=None):
    """Timing rule of the bench: a multi-GPU number is the max over ranks."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64)
    if device is not None:
        t = t.to(device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def allgather_bytes(dist, payload, nbytes, device=None):
    """[bytes of rank 0, bytes of rank 1, ...]: every rank contributes
    `nbytes` bytes (CUDA IPC handles of the exchange areas)."""
    import torch
    raw = bytes(payload)[:nbytes].ljust(nbytes, b"\0")
    mine = torch.tensor(list(raw), dtype=torch.uint8)
    if device is not None:
        mine = mine.to(device)
    outs = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(outs, mine)
    return [bytes(t.cpu().numpy().tobytes()) for t in outs]


def choose_merge(needs_grouping, gh_nslots, part_nparts, key_heap_nslots=0):
    """How the per-GPU states come together before the final Agg:
      'none'     - the state groups by text keys of the session's key heap
                   (long strings travel as session-local words): every rank
                   flushes its own partial rows and PostgreSQL's final Agg
                   merges them, as the reference does (gpupreagg.c:2169-2186)
      'peer'     - no GROUP BY (one record) or a table of at most 64K slots:
                   the ranks push their records into the root's HBM over
                   NVLink peer memory, the root merges (pgs_preagg_merge_peer)
      'exchange' - larger states: the groups are partitioned over the ranks
                   by key hash (pgs_preagg_merge_exchange), every rank
                   flushes its own share"""
    if key_heap_nslots > 0:
        return "none"
    if not needs_grouping:
        return "peer"
    if part_nparts == 0 and gh_nslots <= 65536:
        return "peer"
    return "exchange"


class StateMerge:
    """The merge step of one session per rank.  `ctx` carries rank, world,
    dist (torch.distributed, plumbing only), dev and the NCCL communicator
    `comm` of the C ABI (used by the partitioned exchange)."""

    def __init__(self, lib, sess, ctx, root=0, mode=None):
        from . import _capi
        self.lib, self.sess, self.ctx, self.root = lib, sess, ctx, root
        pm = sess.perfmon()
        self.mode = mode or choose_merge(bool(sess.desc["needs_grouping"]),
                                         pm["gh_nslots"], pm["part_nparts"],
                                         pm.get("key_heap_nslots", 0))
        if self.mode == "peer":
            handle = C.create_string_buffer(64)
            _capi.check(lib.pgs_preagg_peer_setup(sess.handle, ctx.rank, ctx.world, root, handle))
            handles = allgather_bytes(ctx.dist, handle.raw, 64, ctx.dev)
            if ctx.rank != root:
                buf = C.create_string_buffer(handles[root], 64)
                _capi.check(lib.pgs_preagg_peer_attach(sess.handle, buf))
            ctx.dist.barrier()

    def run(self):
        from . import _capi
        if self.mode == "none":
            return
        if self.mode == "peer":
            _capi.check(self.lib.pgs_preagg_merge_peer(self.sess.handle))
        else:
            _capi.check(self.lib.pgs_preagg_merge_exchange(self.sess.handle, self.ctx.comm,
                                                           self.ctx.rank, self.ctx.world))

    def describe(self):
        if self.mode == "none":
            return ("no merge: text keys of the key heap are session-local, every rank returns "
                    "its own partial rows to the final Agg")
        if self.mode == "peer":
            return ("NVLink peer memory: ranks push their state records into the root's HBM "
                    "(gpupreagg_peer_push), the root merges (gpupreagg_peer_pull); no collective")
        return ("hash-partitioned exchange: gpupreagg_export_parts + ncclSend/ncclRecv all-to-all "
                "+ gpupreagg_import; every rank flushes its own share of the groups")

    def trace(self):
        pm = self.sess.perfmon()
        n = max(pm.get("merge_count", 0), 1)
        return {"mode": self.mode, "merges": pm.get("merge_count", 0),
                "merge_kernel_ms_mean": pm.get("merge_kernel_ms", 0.0) / n,
                "merge_exchange_host_ms_mean": pm.get("merge_exchange_ms", 0.0) / n,
                "table_grown": pm.get("num_table_grown", 0)}

    def close(self):
        pass
