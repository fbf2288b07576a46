"""Host-side planning for the multi-GPU build (used by bench.py and the tests; no GPU needed).

The FQB input is grouped by barcode, so a rank's shard is a contiguous range of barcode runs
(SURVEY.md 8e).  These helpers only decide *which* runs each rank takes and move the 128-byte NCCL
unique id between processes; the exchange itself is native (hash10x_b200/csrc/h10x_dist.cuh).
"""
import numpy as np


def plan_shards(run_offsets, world):
    """run_offsets: record offset of every barcode run plus the total (len = runs + 1).
    Returns cut (len world + 1, run indices): rank r takes runs cut[r]..cut[r+1]-1, balanced by record
    count, every rank non-empty when there are at least `world` runs."""
    off = np.asarray(run_offsets, dtype=np.int64)
    runs = off.size - 1
    if runs < world:
        raise ValueError("fewer barcode runs (%d) than ranks (%d)" % (runs, world))
    total = int(off[-1])
    cut = [0]
    for r in range(1, world):
        target = total * r // world
        i = int(np.searchsorted(off, target, side="left"))
        i = max(i, cut[-1] + 1)                 # at least one run per rank
        i = min(i, runs - (world - r))          # leave one run for each later rank
        cut.append(i)
    cut.append(runs)
    return cut


def shard_records(run_offsets, cut, rank):
    off = np.asarray(run_offsets, dtype=np.int64)
    return int(off[cut[rank]]), int(off[cut[rank + 1]])


def share_unique_id(dist, rank, make_id):
    """rank 0 makes the NCCL unique id (make_id()), every rank returns the same 128 bytes."""
    box = [make_id() if rank == 0 else None]
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast_object_list(box, src=0)
    return box[0]


def job_time_and_units(dist, torch, ms_local, units_local, device=None):
    """max over ranks of the device time, sum over ranks of the units processed (bench contract)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(ms_local), float(units_local)
    t = torch.tensor([float(ms_local)], dtype=torch.float64, device=device)
    u = torch.tensor([float(units_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])


def owner_thresholds(k, world, flat=False):
    """hash-range owners (dist_bins step 2): owner o takes hashes in [thr[o], thr[o+1]).  The native function itself
    (h10x_dist_owner_thresholds: host arithmetic, runs without a device), so that models of the exchange cut exactly
    where the library does: at the quantiles of the mosh density, or - flat - into equal widths."""
    import ctypes as C
    from .binding import load_library
    L = load_library()
    L.h10x_dist_owner_thresholds.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p]
    thr = np.zeros(world + 1, np.uint64)
    st = L.h10x_dist_owner_thresholds(k, world, 1 if flat else 0, thr.ctypes.data)
    if st:
        raise ValueError("bad k / world")
    return [int(x) for x in thr]
