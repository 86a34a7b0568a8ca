"""world_size-2 gloo test (CPU): the host side of the sharded path -- segment sharding, per-rank planning, dictionary
export / all_gather / union / import -- must give every rank the same (group x bucket) code space."""
import json
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from lakeside_b200 import synth


def _worker(rank, world, port, paths, full, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from lakeside_b200 import api

    sub, idx = api.shard_request(full, rank, world)
    q = api.Query(json.dumps(sub), aggregates=synth.C2_AGGREGATES, path="dense")
    for i in idx:
        q.add_segment_bytes(open(paths[i], "rb").read())
    q.plan()
    local = q.info
    blobs = [None] * world
    dist.all_gather_object(blobs, q.export_dictionaries())
    q.import_dictionaries(api.union_dictionaries(blobs))
    info = q.info
    json.dump({"local": local, "global": info, "segments": idx, "dicts": [[s.decode() for s in d] for d in api.parse_dictionary_blob(api.union_dictionaries(blobs))]},
              open(os.path.join(out_dir, f"rank{rank}.json"), "w"))
    q.close()
    dist.barrier()
    dist.destroy_process_group()


def test_two_ranks_agree_on_group_space(tmp_path):
    sa = synth.SynthSpec(dataset="metrics", rows=20000, n_names=3, cards=(16, 4, 6, 2))
    sb = synth.SynthSpec(dataset="metrics", rows=20000, n_names=5, cards=(16, 4, 10, 3))
    _, pa = H.dataset("gloo_a", sa, 1)
    _, pb = H.dataset("gloo_b", sb, 1, first_index=100)
    paths = pa + pb  # rank 0 gets segment 0 (spec a), rank 1 gets segment 100 (spec b)
    full = synth.push_down_request(synth.c2_base_expr(), [0, 100], 10000)
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, paths, full, str(tmp_path)), nprocs=2, join=True)
    r0 = json.load(open(tmp_path / "rank0.json"))
    r1 = json.load(open(tmp_path / "rank1.json"))
    assert r0["segments"] == [0] and r1["segments"] == [1]
    # locally the shards see different dictionaries ...
    assert r0["local"]["n_groups"] != r1["local"]["n_groups"]
    # ... after the exchange both index the same space: (5 names + NULL) x (4+1) x (10+1) x (3+1)
    assert r0["global"]["n_groups"] == r1["global"]["n_groups"] == 6 * 5 * 11 * 4
    assert r0["global"]["n_cells"] == r1["global"]["n_cells"] == 6 * 5 * 11 * 4 * 360
    assert r0["dicts"] == r1["dicts"]
    assert r0["dicts"][0] == [f"metric_{i:03d}" for i in range(5)]
    assert [k["dict"] for k in r0["global"]["keys"]] == [5, 4, 10, 3]
