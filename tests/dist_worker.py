"""Worker of the multi-GPU parity test: run as `python -m torch.distributed.run --nproc-per-node N
tests/dist_worker.py` (or plainly, as a single rank).  Each rank builds its barcode-range shard with
h10x_gpu_build_device_dist; rank 0 reassembles the global index from the per-rank pieces and compares it
strictly (bin ids, values, depths, ClusterHash lists, hash->code lists, hashIndex) with the CPU oracle
run on the whole data set."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import hashfile  # noqa: E402
import hash10x_b200  # noqa: E402
from oracle import orc  # noqa: E402


def main():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    torch.cuda.set_device(local)
    B = 21
    for case, pkw in enumerate([dict(seed=41, n_barcodes=37, pairs_min=1, pairs_max=150),
                                dict(seed=42, n_barcodes=200, pairs_min=20, pairs_max=300, read_len=160),
                                dict(seed=43, n_barcodes=world, pairs_min=5, pairs_max=9)]):
        p = orc.synth_params(**pkw)
        n, off = orc.synth_layout(p)
        nb = p.nBarcodes
        cut = [int(round(nb * r / world)) for r in range(world + 1)]
        r0, r1 = int(off[cut[rank]]), int(off[cut[rank + 1]])
        recs = orc.synth_fqb(p, r0, r1)
        fqb = torch.from_numpy(recs.view(np.int32).reshape(-1).copy()).cuda()
        g = hash10x_b200.Hash10xGPU(B=B, device=local)
        idb = [hash10x_b200.Hash10xGPU.dist_unique_id() if rank == 0 else None]
        if world > 1:
            dist.broadcast_object_list(idb, src=0)
        g.dist_init(rank, world, idb[0])
        if os.environ.get("H10X_DIST_FAIL"):
            # a rank-local failure inside the bin exchange: EVERY rank must come back with that error (nobody left waiting
            # in a collective), and the same contexts must then be able to build again
            try:
                g.build_device_dist(fqb.data_ptr(), recs.shape[0])
                raise SystemExit("rank %d: the injected failure did not surface" % rank)
            except hash10x_b200.H10xError as e:
                assert e.code == 4, (e.code, str(e))            # H10X_ERR_NOMEM on every rank
            del os.environ["H10X_DIST_FAIL"]
            if world > 1:
                dist.barrier()
            print("rank %d: injected failure surfaced everywhere" % rank, flush=True)
        g.build_device_dist(fqb.data_ptr(), recs.shape[0])
        ix = g.download()
        info = g.dist_info()
        # the rows that follow the build, per rank on its own blocks: --hashDepthRange and --cluster need every bin's depth
        # and its barcodes on ALL ranks (hash10x.c:528-539, 738-868)
        g.dist_global_codes()
        _within, goff, good = g.depth_range(2, 40)
        cclus, nsub, ptm, _ms = g.cluster(0, 0, 2)
        gcoff, gcodes = g.download_codes()
        piece = dict(rank=rank, blkNRead=ix.blkNRead, blkNHash=ix.blkNHash, clus=ix.clus, info=info,
                     goff=goff, good=good, cclus=cclus, nsub=nsub, ptm=ptm, gcoff=gcoff, gcodes=gcodes,
                     hashNumber=ix.hashNumber, hashValue=ix.hashValue if rank == 0 else None,
                     hashDepth=ix.hashDepth if rank == 0 else None, hashIndex=ix.hashIndex if rank == 0 else None)
        pieces = [None] * world
        if world > 1:
            dist.gather_object(piece, pieces if rank == 0 else None, dst=0)
        else:
            pieces = [piece]
        if rank == 0:
            want = orc.build(orc.synth_fqb(p), B=B)
            assert want.status == 0
            hn = pieces[0]["hashNumber"]
            assert hn == want.hashNumber, (hn, want.hashNumber)
            assert np.array_equal(pieces[0]["hashValue"], want.hashValue)
            assert np.array_equal(pieces[0]["hashDepth"], want.hashDepth)
            assert np.array_equal(pieces[0]["hashIndex"], want.hashIndex)
            nread = np.concatenate([[0]] + [pc["blkNRead"][1:] for pc in pieces]).astype(np.uint32)
            nhash = np.concatenate([[0]] + [pc["blkNHash"][1:] for pc in pieces]).astype(np.uint32)
            clus = np.concatenate([pc["clus"] for pc in pieces])
            assert np.array_equal(nread, want.blkNRead) and np.array_equal(nhash, want.blkNHash)
            assert np.array_equal(clus, want.clus)
            base = 0
            for pc in pieces:
                assert pc["info"]["blockBase"] == base and pc["info"]["nBlocksGlobal"] == want.nBlocksMax - 1
                base += pc["blkNRead"].size - 1
            # hash->code lists: concatenation over ranks of each bin's local list
            lists = [[] for _ in range(hn)]
            for pc in pieces:
                inf = pc["info"]
                for j, b in enumerate(inf["localBinId"]):
                    lists[int(b)].append(inf["localCodes"][int(inf["localCodeOff"][j]):int(inf["localCodeOff"][j + 1])])
            for x in range(1, hn):
                got = np.concatenate(lists[x]) if lists[x] else np.zeros(0, np.uint32)
                assert np.array_equal(got, want.codes[int(want.codeOff[x]):int(want.codeOff[x + 1])]), x
            # every rank holds the whole hash->code CSR; good lists / sub-clusters of the ranks' blocks, put end to end, are
            # what the oracle gets for the whole data set
            _w, wgoff, wgood = orc.good_hashes(want, 2, 40)
            wclus, wnsub, wptm = orc.cluster(want, wgoff, wgood, 0, 0, 2)
            for pc in pieces:
                assert np.array_equal(pc["gcoff"], want.codeOff) and np.array_equal(pc["gcodes"], want.codes), pc["rank"]
            ggood = np.concatenate([pc["good"] for pc in pieces])
            gsizes = np.concatenate([[0]] + [np.diff(pc["goff"].astype(np.int64))[1:] for pc in pieces])
            assert np.array_equal(ggood, wgood) and np.array_equal(gsizes, np.diff(wgoff.astype(np.int64)))
            assert np.array_equal(np.concatenate([pc["cclus"] for pc in pieces]), wclus)
            assert np.array_equal(np.concatenate([[0]] + [pc["nsub"][1:] for pc in pieces]).astype(np.uint32), wnsub)
            gptm = np.concatenate([[0.0]] + [pc["ptm"][1:] for pc in pieces])
            assert np.array_equal(gptm.view(np.uint64), wptm.view(np.uint64))
            print("dist case %d ok: %d ranks, %d bins, %d hashes, %d good, %d sub-clusters" %
                  (case, world, hn - 1, clus.size, wgood.size, int(wnsub.sum())), flush=True)
        g.close()
        if world > 1:
            dist.barrier()
    # a failure on one rank only (after the runs were agreed on) must come back as an error on every rank, not as a hang
    os.environ["H10X_TEST_FAIL_RANK"] = str(world - 1)
    p = orc.synth_params(seed=44, n_barcodes=4 * world, pairs_min=5, pairs_max=40)
    n, off = orc.synth_layout(p)
    cut = [int(round(p.nBarcodes * r / world)) for r in range(world + 1)]
    recs = orc.synth_fqb(p, int(off[cut[rank]]), int(off[cut[rank + 1]]))
    fqb = torch.from_numpy(recs.view(np.int32).reshape(-1).copy()).cuda()
    g = hash10x_b200.Hash10xGPU(B=B, device=local)
    idb = [hash10x_b200.Hash10xGPU.dist_unique_id() if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(idb, src=0)
    g.dist_init(rank, world, idb[0])
    try:
        g.build_device_dist(fqb.data_ptr(), recs.shape[0])
        failed = False
    except hash10x_b200.H10xError as e:
        failed = e.code == 8
    del os.environ["H10X_TEST_FAIL_RANK"]
    assert failed, "rank %d: the forced failure of rank %d did not arrive" % (rank, world - 1)
    g.build_device_dist(fqb.data_ptr(), recs.shape[0])          # the context and the communicator are still usable
    g.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        print("DIST_PARITY_OK world=%d" % world, flush=True)


if __name__ == "__main__":
    main()
