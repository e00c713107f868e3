"""Oracle-exact parity for COHORT mode (csrc/cohort.cu: several users' trees share every forward; SURVEY 8f-3 "per-user
results unchanged by batching").

The fp32 exact-parity forward (csrc/forward_f32.cu) runs inside the cohort scheduler, so the kernels the benchmark times
-- cohort_begin / cohort_build_batch / cohort_select / cohort_verify / kv_gather_cohort / cohort_results and the packing
logic of bssd_batch -- are replayed against fixtures written by the UNMODIFIED reference (tools/make_golden*.py):

  * tests/golden/bssd_strict_users.json: 64 Beauty + 64 Games users, K=10 N=40 gamma=3, strict / positional, through
    cohorts of 8 and of 16 users in flight;
  * tests/golden/bssd_strict.json (hf_fp32 stack): the K x N x gamma grid, each configuration's users in one cohort.

Ranked lists, accepted lengths per round and n_run must be the reference's; scores within 1e-3 relative.  The one
admissible exception is the fp32 exact-tie swap of tests/test_zz_gpu_fp32_users.py (two beams whose scores coincide to
1e-4 absolute), at most 3 users in 64.  A last test feeds the cohort's ranked lists to the metrics mirror and requires
Recall@K / NDCG@K identical to the values the reference's computeTopNAccuracy produced (code/utils.py:215-271)."""
import collections

import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, golden, lists_match, stack_weights

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def fp32_models():
    from atspeed_b200.engine import DeviceModel, ModelSpec
    cache = {}

    def get(ds_name, which):
        if (ds_name, which) not in cache:
            sh, W = stack_weights("hf_fp32", ds_name, which)
            spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
            cache[(ds_name, which)] = DeviceModel(spec, W, "cuda", dtype=torch.float32)
        return cache[(ds_name, which)]

    return get


def cohort_session(fp32_models, ds_name, kind, draft, K, N, max_users):
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceTrie, Session
    ds = dataset(ds_name)
    csr = compile_constraint(constraint_fn(ds_name, kind), ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
    trie = DeviceTrie(csr, torch.device("cuda"))
    return ds, Session(fp32_models(ds_name, "target"), fp32_models(ds_name, draft), trie, K, N, 4, max_users=max_users)


def check_against_golden(got, case, tally):
    items, scores = got["tokens"].tolist(), got["scores"]
    if items == case["bssd"]["items"]:
        tally["exact"] += 1
        want = np.asarray(case["bssd"]["scores"], dtype=np.float64)
        tally["rel"] = max(tally["rel"], float(np.max(np.abs(scores - want) / np.abs(want))))
    else:
        ok, _, msg = lists_match(items, scores, case["bssd"]["items"], case["bssd"]["scores"], 1e-4)
        assert ok, f"user {case['user']}: {msg}"
    accept = case["accept_steps"] if "accept_steps" in case else [r["n_matches"] for r in case["rounds"]]
    tally["steps"] += int(got["accept_steps"] == accept and got["n_run"] == case["n_run"])
    tally["n"] += 1


@pytest.mark.parametrize("max_users", [8, 16])
@pytest.mark.parametrize("kind", ["strict", "positional"])
@pytest.mark.parametrize("ds_name", ["beauty", "games"])
def test_cohort_fp32_matches_reference_on_golden_users(fp32_models, ds_name, kind, max_users):
    cases = [c for c in golden("bssd_strict_users.json")["cases"] if c["dataset"] == ds_name and c["constraint"] == kind]
    assert len(cases) == 32 and all(c["K"] == 10 and c["N"] == 40 and c["gamma"] == 3 for c in cases)
    ds, sess = cohort_session(fp32_models, ds_name, kind, "correlated", 10, 40, max_users)
    got = sess.bssd_batch([ds.prompt_ids(c["user"]) for c in cases], 3)
    tally = collections.Counter(rel=0.0)
    for g, c in zip(got, cases):
        check_against_golden(g, c, tally)
    print(f"cohort fp32 {ds_name}/{kind}, {max_users} users in flight: {tally['exact']}/32 ranked lists identical to the "
          f"reference, {tally['steps']}/32 accepted-length sequences identical, max relative score error {tally['rel']:.2e}")
    assert tally["rel"] < 1e-3
    assert tally["exact"] >= 30 and tally["steps"] >= 30, dict(tally)


def _common_prefix(prompts):
    n = min(len(p) for p in prompts) - 1
    for p in prompts[1:]:
        k = 0
        while k < n and p[k] == prompts[0][k]:
            k += 1
        n = k
    return list(prompts[0][:n])


@pytest.mark.parametrize("ds_name,kind", [("beauty", "strict"), ("games", "positional")])
def test_cohort_fp32_with_shared_prompt_prefix_matches_reference(fp32_models, ds_name, kind):
    """atspeed_session_set_shared_prefix: the K/V rows of the prompts' common opening tokens are computed once per session and
    copied into every user's caches; the users' forwards skip those tokens.  In the fp32 parity mode nothing may change: ranked
    lists, accepted lengths and n_run are still the reference's."""
    cases = [c for c in golden("bssd_strict_users.json")["cases"] if c["dataset"] == ds_name and c["constraint"] == kind]
    ds, sess = cohort_session(fp32_models, ds_name, kind, "correlated", 10, 40, 16)
    prompts = [ds.prompt_ids(c["user"]) for c in cases]
    prefix = _common_prefix(prompts)
    assert len(prefix) >= 8, "the datasets' prompts share their instruction template"
    assert sess.set_shared_prefix(prefix) == len(prefix)
    got = sess.bssd_batch(prompts, 3)
    tally = collections.Counter(rel=0.0)
    for g, c in zip(got, cases):
        check_against_golden(g, c, tally)
    print(f"cohort fp32 {ds_name}/{kind} with a shared prefix of {len(prefix)} tokens: {tally['exact']}/32 ranked lists identical to "
          f"the reference, {tally['steps']}/32 accepted-length sequences identical, max relative score error {tally['rel']:.2e}")
    assert tally["rel"] < 1e-3
    assert tally["exact"] >= 30 and tally["steps"] >= 30, dict(tally)
    # a prompt that does not start with the prefix is refused, not silently mis-evaluated
    from atspeed_b200._lib import AtSpeedError
    bad = list(prompts[0])
    bad[3] = bad[3] + 1 if bad[3] + 1 < ds.vocab_size else bad[3] - 1
    with pytest.raises(AtSpeedError, match="shared"):
        sess.bssd_batch([bad, prompts[1]], 3)
    # and switching it off restores the plain path
    assert sess.set_shared_prefix([]) == 0
    again = sess.bssd_batch(prompts[:4], 3)
    for g, c in zip(again, cases[:4]):
        check_against_golden(g, c, collections.Counter(rel=0.0))


def test_cohort_fp32_matches_reference_on_the_config_grid(fp32_models):
    """Every (dataset, constraint, draft, K, N, gamma) configuration of the hf_fp32 golden grid: its users as one cohort."""
    groups = collections.defaultdict(list)
    for c in golden()["cases"]:
        if c["stack"] == "hf_fp32":
            groups[(c["dataset"], c["constraint"], c["draft"], c["K"], c["N"], c["gamma"])].append(c)
    tally = collections.Counter(rel=0.0)
    for (ds_name, kind, draft, K, N, gamma), cases in sorted(groups.items()):
        ds, sess = cohort_session(fp32_models, ds_name, kind, draft, K, N, 4)
        got = sess.bssd_batch([ds.prompt_ids(c["user"]) for c in cases], gamma)
        for g, c in zip(got, cases):
            check_against_golden(g, c, tally)
        del sess
    print(f"cohort fp32 over the golden grid: {tally['exact']}/{tally['n']} ranked lists identical, {tally['steps']}/{tally['n']} "
          f"accepted-length sequences identical, max relative score error {tally['rel']:.2e}")
    assert tally["n"] == 184 and tally["rel"] < 1e-3
    assert tally["exact"] >= tally["n"] - 4 and tally["steps"] >= tally["n"] - 4, dict(tally)


@pytest.mark.parametrize("ds_name", ["beauty", "games"])
def test_recall_ndcg_of_gpu_ranked_lists_equal_the_reference(fp32_models, ds_name):
    """Ranked lists produced on the GPU (cohorts of 16, fp32 mode) -> runner.gather_records -> runner.evaluate's decode path
    -> computeTopNAccuracy must give (i) the metrics of the reference's own recorded lists and (ii), with the fixture's
    planted hits, exactly the four vectors the reference's computeTopNAccuracy returned (tests/golden/...users.json)."""
    from atspeed_b200.metrics import computeTopNAccuracy
    from atspeed_b200.runner import UserRecords, evaluate, gather_records
    g = golden("bssd_strict_users.json")
    cases = [c for c in g["cases"] if c["dataset"] == ds_name]
    ds = dataset(ds_name)
    rec = UserRecords.empty(len(cases), 10, 4)
    for kind in ("strict", "positional"):
        idx = [i for i, c in enumerate(cases) if c["constraint"] == kind]
        _, sess = cohort_session(fp32_models, ds_name, kind, "correlated", 10, 40, 16)
        outs = sess.bssd_batch([ds.prompt_ids(cases[i]["user"]) for i in idx], 3)
        for i, o in zip(idx, outs):
            rec.users[i] = cases[i]["user"]
            rec.items[i] = o["tokens"]
            rec.scores[i] = o["scores"]
            rec.meta[i] = (o["n_run"], o["total_accept_steps"], 0, 10)
    order = np.argsort(rec.users, kind="stable")
    full = gather_records(rec, 10, 4, per_rank=len(cases))
    assert full.users.tolist() == rec.users[order].tolist()
    topN = g["metrics"]["topN"]
    m = evaluate(ds, full, topN)
    by_user = {c["user"]: c for c in cases}
    gts = [ds.ground_truth_strings(int(u)) for u in full.users]
    ref_preds = [ds.decode_items(by_user[int(u)]["bssd"]["items"]) for u in full.users]
    want = computeTopNAccuracy(gts, ref_preds, topN)
    assert (m["precision"], m["recall"], m["ndcg"], m["mrr"]) == tuple(want), "metrics of GPU lists != metrics of reference lists"
    # the fixture's planted variant, in the fixture's user order
    gts, preds = [], []
    for i, c in enumerate(cases):
        gt = ds.ground_truth_strings(c["user"])
        names = ds.decode_items(rec.items[i])
        if i % 3 != 2:
            names[i % 10] = gt[0]
        gts.append(gt)
        preds.append(names)
    got = computeTopNAccuracy(gts, preds, topN)
    assert [list(x) for x in got] == g["metrics"]["values"][ds_name]
