"""Host-side plumbing for one-process-per-GPU runs (torchrun): which pairs a rank solves, the
max-over-ranks of a device time, and the final gather of per-pair results on rank 0.
Pairs are independent, so there is no collective on the data path; these
helpers are the only distributed code and work with the nccl (GPU) and gloo (CPU tests) back-ends."""
import numpy as np


def shard_indices(n_items, rank, world):
    """Round-robin pair -> rank map (pair i goes to rank i mod world)."""
    return list(range(rank, n_items, world))


def _dist():
    import torch.distributed as dist
    return dist if dist.is_available() and dist.is_initialized() else None


class WorkQueue:
    """Dynamic work sharing between the ranks of a torchrun job: an atomic counter in the process group's key-value
    store (TCPStore.add), so that a rank that drew short solves (outer and CG iteration counts are data dependent)
    takes the next item instead of idling -- the multi-process counterpart of the atomic work queue inside
    foto_solve_batch.  No collective and no data on the path: only item indices travel.  Without a process group
    it counts locally.  `name` must be unique per queue (counters are never reset)."""

    def __init__(self, name, n_items, chunk=1):
        self.n, self.chunk, self.key = int(n_items), max(int(chunk), 1), f"foto_b200/wq/{name}"
        self._local, self._have = 0, []
        d = _dist()
        self._store = None
        if d is not None:
            from torch.distributed import distributed_c10d as c10d
            self._store = c10d._get_default_store()

    def next(self):
        """Next unclaimed item index, or None when the queue is empty."""
        if not self._have:
            if self._store is None:
                first = self._local; self._local += self.chunk
            else:
                first = self._store.add(self.key, self.chunk) - self.chunk
            self._have = [i for i in range(first, first + self.chunk) if i < self.n][::-1]
            if not self._have:
                return None
        return self._have.pop()


def gather_objects(obj):
    """List of every rank's picklable `obj` on every rank (single-process: [obj])."""
    d = _dist()
    if d is None:
        return [obj]
    out = [None] * d.get_world_size()
    d.all_gather_object(out, obj)
    return out


def barrier():
    d = _dist()
    if d is not None:
        d.barrier()


def max_over_ranks(values, device=None):
    """Element-wise max of a list of floats over all ranks (identity without a process group)."""
    d = _dist()
    if d is None:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    d.all_reduce(t, op=d.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def gather_rows(local_rows, local_indices, n_items, row_len, device=None):
    """Assemble the (n_items, row_len) result on rank 0 from every rank's rows; None elsewhere."""
    d = _dist()
    local_rows = np.asarray(local_rows, dtype=np.float64).reshape(len(local_indices), row_len)
    if d is None:
        out = np.empty((n_items, row_len))
        out[local_indices] = local_rows
        return out
    import torch
    world, rank = d.get_world_size(), d.get_rank()
    per = (n_items + world - 1) // world
    buf = torch.zeros((per, row_len), dtype=torch.float64, device=device)
    if len(local_indices):
        buf[:len(local_indices)] = torch.from_numpy(local_rows).to(buf.device)
    bufs = [torch.empty_like(buf) for _ in range(world)] if rank == 0 else None
    d.gather(buf, bufs, dst=0)
    if rank != 0:
        return None
    out = np.empty((n_items, row_len))
    for r in range(world):
        idx = shard_indices(n_items, r, world)
        out[idx] = bufs[r][:len(idx)].cpu().numpy()
    return out
