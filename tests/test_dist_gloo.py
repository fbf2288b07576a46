"""CPU tests (gloo, world_size 2) of the host-side multi-GPU logic: shard planning, unique-id sharing,
max/sum reduction of the bench contract, and a numpy MODEL of the exchange protocol that
hash10x_b200/csrc/h10x_dist.cuh implements with NCCL (hash-range owners, depth = sum, first block = min,
bin id = 1 + new hashes of earlier blocks + new hashes of earlier owners in the same block + rank by hash).
The model's global ids must equal the oracle's on the whole data set."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p_ in (ROOT, os.path.join(ROOT, "tests")):
    if p_ not in sys.path:
        sys.path.insert(0, p_)

from hash10x_b200 import shard  # noqa: E402


def test_plan_shards_properties():
    rng = np.random.default_rng(1)
    for runs in (2, 3, 17, 1000):
        sizes = rng.integers(1, 500, runs)
        off = np.r_[0, np.cumsum(sizes)]
        for world in (1, 2, 4, 8):
            if runs < world:
                with pytest.raises(ValueError):
                    shard.plan_shards(off, world)
                continue
            cut = shard.plan_shards(off, world)
            assert cut[0] == 0 and cut[-1] == runs and len(cut) == world + 1
            assert all(cut[i] < cut[i + 1] for i in range(world))
            spans = [shard.shard_records(off, cut, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == off[-1]
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            if runs >= 100 * world:
                share = [(b - a) / off[-1] for a, b in spans]
                assert max(share) < 1.2 / world


def test_owner_thresholds_cover_the_hash_space():
    for k in (13, 21, 31):
        for world in (1, 2, 3, 8):
            for flat in (False, True):
                thr = shard.owner_thresholds(k, world, flat)
                assert thr[0] == 0 and thr[-1] == 1 << (2 * k) and all(a < b for a, b in zip(thr, thr[1:]))
            # the quantile cut gives the low owners narrower ranges (the density of min (hash, hashRC) falls with the hash)
            q = shard.owner_thresholds(k, world)
            if world > 1:
                assert q[1] - q[0] < q[-1] - q[-2]


def _alltoallv(send, recv, rank, world):
    """send/recv pairs in one batch, as the native code does with ncclSend/ncclRecv in a group
    (gloo has no all_to_all)"""
    recv[rank].copy_(send[rank])
    ops = []
    for peer in range(world):
        if peer == rank:
            continue
        if send[peer].numel():
            ops.append(dist.P2POp(dist.isend, send[peer].contiguous(), peer))
        if recv[peer].numel():
            ops.append(dist.P2POp(dist.irecv, recv[peer], peer))
    if ops:
        for w in dist.batch_isend_irecv(ops):
            w.wait()


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import orc
        # --- helpers of the bench contract ---
        idb = shard.share_unique_id(dist, rank, lambda: bytes(range(128)))
        assert idb == bytes(range(128))
        ms, units = shard.job_time_and_units(dist, torch, 10.0 + rank, 100 * (rank + 1))
        assert ms == 10.0 + world - 1 and units == 100 * world * (world + 1) / 2

        # --- the exchange protocol, modelled in numpy over gloo ---
        k, B = 21, 21
        p = orc.synth_params(seed=51, n_barcodes=41, pairs_min=1, pairs_max=120)
        n, off = orc.synth_layout(p)
        full = orc.synth_fqb(p)
        cut = shard.plan_shards(off, world)
        r0, r1 = shard.shard_records(off, cut, rank)
        mine = full[r0:r1]
        last = rank == world - 1
        if not last:                      # only the GLOBALLY last run stays unhashed: add a sentinel run
            sentinel = full[r1:r1 + 1].copy()
            mine = np.concatenate([mine, sentinel])
        loc = orc.build(mine, B=B)
        assert loc.status == 0
        n_blk = cut[rank + 1] - cut[rank]
        blk_base = cut[rank]
        hv = loc.hashValue[1:]
        depth = loc.hashDepth[1:]
        first = loc.codes[loc.codeOff[1:-1].astype(np.int64)].astype(np.int64) + blk_base   # first (global) block
        order = np.argsort(hv, kind="stable")
        hv, depth, first = hv[order], depth[order].astype(np.int64), first[order]
        thr = np.array(shard.owner_thresholds(k, world), dtype=np.uint64)
        send_off = np.searchsorted(hv, thr, side="left")
        send = [torch.from_numpy(np.stack([hv[send_off[o]:send_off[o + 1]].astype(np.int64),
                                           depth[send_off[o]:send_off[o + 1]],
                                           first[send_off[o]:send_off[o + 1]]], 1).copy()) for o in range(world)]
        cnt = torch.tensor([t.shape[0] for t in send], dtype=torch.int64)
        cnts = [torch.zeros(world, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(cnts, cnt)
        recv = [torch.zeros((int(cnts[src][rank]), 3), dtype=torch.int64) for src in range(world)]
        _alltoallv(send, recv, rank, world)
        got = np.concatenate([t.numpy() for t in recv]) if world > 1 else recv[0].numpy()
        src_of = np.concatenate([np.full(t.shape[0], i) for i, t in enumerate(recv)])
        # owner merge
        o = np.argsort(got[:, 0].astype(np.uint64), kind="stable")
        gh = got[o, 0].astype(np.uint64)
        head = np.r_[True, gh[1:] != gh[:-1]] if gh.size else np.zeros(0, bool)
        seg = np.cumsum(head) - 1
        n_seg = int(seg[-1]) + 1 if gh.size else 0
        g_hash = gh[head]
        g_depth = np.bincount(seg, weights=got[o, 1], minlength=n_seg).astype(np.int64)
        g_first = np.full(n_seg, np.iinfo(np.int64).max)
        np.minimum.at(g_first, seg, got[o, 2])
        n_b2 = len(off) - 1 + 2
        new_cnt = torch.from_numpy(np.bincount(g_first, minlength=n_b2).astype(np.int64))
        mats = [torch.zeros(n_b2, dtype=torch.int64) for _ in range(world)]
        dist.all_gather(mats, new_cnt)
        mat = np.stack([m.numpy() for m in mats])
        col = mat.sum(0)
        prefix = np.r_[0, np.cumsum(col)[:-1]]
        below = mat[:rank].sum(0) if rank else np.zeros(n_b2, np.int64)
        so = np.argsort(g_first, kind="stable")          # g_hash is ascending: stable sort = (first, hash)
        sf = g_first[so]
        ghead = np.r_[True, sf[1:] != sf[:-1]] if sf.size else np.zeros(0, bool)
        gstart = np.maximum.accumulate(np.where(ghead, np.arange(sf.size), 0)) if sf.size else sf
        g_id = np.zeros(n_seg, np.int64)
        g_id[so] = 1 + prefix[sf] + below[sf] + (np.arange(sf.size) - gstart)
        ans = np.zeros(got.shape[0], np.int64)
        ans[o] = g_id[seg]
        back_send = [torch.from_numpy(ans[src_of == src].copy()) for src in range(world)]
        back_recv = [torch.zeros(int(cnt[o_]), dtype=torch.int64) for o_ in range(world)]
        _alltoallv(back_send, back_recv, rank, world)
        my_ids = np.concatenate([t.numpy() for t in back_recv])
        # --- compare with the oracle on the whole data set ---
        want = orc.build(full, B=B)
        ref = dict(zip(want.hashValue[1:].tolist(), range(1, want.hashNumber)))
        assert int(col.sum()) == want.hashNumber - 1
        assert [ref[int(h)] for h in hv] == my_ids.tolist()
        for h, d_, i in zip(g_hash.tolist(), g_depth.tolist(), g_id.tolist()):
            assert want.hashValue[i] == h and want.hashDepth[i] == d_
        q.put((rank, "ok", n_blk))
    except Exception as e:  # pragma: no cover
        import traceback
        q.put((rank, "fail: " + repr(e) + traceback.format_exc(), 0))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2])
def test_protocol_model_over_gloo(world):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for pr in procs:
        pr.start()
    res = [q.get(timeout=300) for _ in range(world)]
    for pr in procs:
        pr.join(timeout=60)
    assert all(r[1] == "ok" for r in res), res
    assert sum(r[2] for r in res) == 41


def test_owner_merge_tiling_model():
    """The owner's merge of its received runs (h10x_dist.cuh "owner merge without a sort"), restated in numpy: every S-th
    element of every run is a splitter candidate, every NR-th sorted candidate a tile boundary, a tile takes from each run
    the elements in [B_t, B_t+1).  Checked here: the tiles partition every run, no tile reaches 3 NR S elements whatever
    the runs look like (disjoint ranges, one run dense and the others sparse, shared values), and merging tile after tile
    with ties in source order is the stable sort of the concatenation - which is what the sort path computes."""
    rng = np.random.default_rng(3)
    for nr, S, shape in [(2, 16, "uniform"), (8, 8, "uniform"), (8, 8, "disjoint"), (5, 4, "skewed"), (8, 8, "shared"), (3, 8, "empty")]:
        runs = []
        for r in range(nr):
            n = int(rng.integers(200, 2000))
            if shape == "disjoint":
                v = rng.choice(np.arange(r * 100_000, (r + 1) * 100_000), n, replace=False)
            elif shape == "skewed":
                v = rng.choice(np.arange(0, 1_000_000), n * (20 if r == 0 else 1) // (1 if r == 0 else 4), replace=False)
            elif shape == "shared":
                v = rng.choice(np.arange(0, 3000), min(n, 2500), replace=False)     # most values held by several runs
            elif shape == "empty" and r == 1:
                v = np.zeros(0, np.int64)
            else:
                v = rng.choice(np.arange(0, 1_000_000), n, replace=False)
            runs.append(np.sort(v.astype(np.int64)))                                 # rank-distinct: no value twice in a run
        cand = np.sort(np.concatenate([run[S::S][:(run.size - 1) // S] if run.size else run for run in runs]))
        n_tiles = cand.size // nr + 1
        bounds = [None] + [int(cand[t * nr - 1]) for t in range(1, n_tiles)] + [None]
        start = [[0 if bounds[t] is None and t == 0 else (run.size if bounds[t] is None else int(np.searchsorted(run, bounds[t], "left")))
                  for run in runs] for t in range(n_tiles + 1)]
        merged_val, merged_src = [], []
        for t in range(n_tiles):
            pieces = [runs[r][start[t][r]:start[t + 1][r]] for r in range(nr)]
            size = sum(p.size for p in pieces)
            assert size < 3 * nr * S, (shape, nr, S, size)
            vals = np.concatenate(pieces)
            src = np.concatenate([np.full(p.size, r) for r, p in enumerate(pieces)])
            order = np.lexsort((src, vals))                                          # ties: the lower source rank first
            merged_val.append(vals[order])
            merged_src.append(src[order])
        mv, ms = np.concatenate(merged_val), np.concatenate(merged_src)
        allv = np.concatenate(runs)
        alls = np.concatenate([np.full(run.size, r) for r, run in enumerate(runs)])
        order = np.lexsort((alls, allv))
        assert np.array_equal(mv, allv[order]) and np.array_equal(ms, alls[order]), shape
