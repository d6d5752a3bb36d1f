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
