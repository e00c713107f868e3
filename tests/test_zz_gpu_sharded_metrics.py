"""Metrics end to end across GPUs (north star: "Recall@K and NDCG@K must be identical").

Two ranks, one GPU each (skipped below two devices): the 64 golden users of a dataset are sharded r, r+W, ... exactly as
atspeed_b200/runner.py and bench.py do, every rank runs its slice through the CUDA path (fp32 exact-parity mode, cohorts of 16
users in flight), ONE NCCL all-gather of the fixed-stride records (runner.gather_records), then every rank computes
Recall/NDCG/MRR/Precision with the mirror of the reference's computeTopNAccuracy (code/utils.py:215-271).  Required:
  * both ranks hold identical ranked lists and identical metrics;
  * the metrics equal those of the reference's own recorded lists (tests/golden/bssd_strict_users.json);
  * with the fixture's planted hits they equal the four vectors the REFERENCE's computeTopNAccuracy returned."""
import os
import socket
import sys

import numpy as np
import pytest
import torch

from _common import ROOT, dataset, golden

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out_dir, ds_name):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from _common import constraint_fn, dataset as _ds, golden as _golden, stack_weights
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceModel, DeviceTrie, ModelSpec, Session
    from atspeed_b200.metrics import computeTopNAccuracy
    from atspeed_b200.runner import UserRecords, evaluate, gather_records, shard_users
    g = _golden("bssd_strict_users.json")
    cases = [c for c in g["cases"] if c["dataset"] == ds_name]
    ds = _ds(ds_name)
    models = {}
    for which in ("target", "correlated"):
        sh, W = stack_weights("hf_fp32", ds_name, which)
        spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
        models[which] = DeviceModel(spec, W, dev, dtype=torch.float32)
    mine = shard_users(list(range(len(cases))), rank, world)             # indices into the fixture's user list
    rec = UserRecords.empty(len(mine), 10, 4)
    for kind in ("strict", "positional"):
        idx = [j for j, i in enumerate(mine) if cases[i]["constraint"] == kind]
        if not idx:
            continue
        csr = compile_constraint(constraint_fn(ds_name, kind), ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
        sess = Session(models["target"], models["correlated"], DeviceTrie(csr, dev), 10, 40, 4, max_users=16)
        outs = sess.bssd_batch([ds.prompt_ids(cases[mine[j]]["user"]) for j in idx], 3)
        for j, o in zip(idx, outs):
            rec.users[j] = mine[j]                       # record key = position in the fixture (users are not unique keys)
            rec.items[j], rec.scores[j] = o["tokens"], o["scores"]
            rec.meta[j] = (o["n_run"], o["total_accept_steps"], 0, 10)
    full = gather_records(rec, 10, 4, per_rank=-(-len(cases) // world), device=dev)
    assert full.users.tolist() == list(range(len(cases)))
    topN = g["metrics"]["topN"]
    gts = [ds.ground_truth_strings(c["user"]) for c in cases]
    preds = [ds.decode_items(full.items[i]) for i in range(len(cases))]
    got = computeTopNAccuracy(gts, preds, topN)
    want = computeTopNAccuracy(gts, [ds.decode_items(c["bssd"]["items"]) for c in cases], topN)
    planted = []
    for i, (gt, names) in enumerate(zip(gts, preds)):
        names = list(names)
        if i % 3 != 2:
            names[i % 10] = gt[0]
        planted.append(names)
    got_planted = computeTopNAccuracy(gts, planted, topN)
    np.savez(os.path.join(out_dir, f"r{rank}.npz"), items=full.items, scores=full.scores, got=np.asarray(got), want=np.asarray(want),
             planted=np.asarray(got_planted), exact=np.asarray([full.items[i].tolist() == c["bssd"]["items"] for i, c in enumerate(cases)]))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs (gpurun --gpus 2)")
@pytest.mark.parametrize("ds_name", ["beauty", "games"])
def test_recall_ndcg_identical_on_two_gpus_and_equal_to_the_reference(ds_name, tmp_path):
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), ds_name), nprocs=2, join=True)
    a, b = np.load(tmp_path / "r0.npz"), np.load(tmp_path / "r1.npz")
    for k in ("items", "scores", "got", "planted"):
        np.testing.assert_array_equal(a[k], b[k])                 # identical ranked lists and metrics on every rank
    np.testing.assert_array_equal(a["got"], a["want"])           # == metrics of the reference's recorded lists
    assert a["planted"].tolist() == golden("bssd_strict_users.json")["metrics"]["values"][ds_name]
    print(f"{ds_name}: {int(a['exact'].sum())}/64 ranked lists identical to the reference's; Recall@10 {a['got'][1][-1]}, "
          f"NDCG@10 {a['got'][2][-1]} on both ranks")
    assert int(a["exact"].sum()) >= 61
