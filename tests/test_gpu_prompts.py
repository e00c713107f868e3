"""Device-side prompt builder (csrc/prompt.cu) against the host tokenisation it replaces
(atspeed_b200/prompts.py RecDataset.prompt_ids, the synthetic stand-in of reference code/data.py:232-263 +
code/collator.py:50-75): bit-exact token ids for users of every history length, and a search fed with device-built prompts
returns what the same search returns for host-built ones."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset, stack_weights

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["beauty", "games"])
def test_device_prompts_equal_host_prompts(name):
    from atspeed_b200.prompts import DevicePromptBuilder
    ds = dataset(name)
    b = DevicePromptBuilder(ds, "cuda")
    hl = np.diff(ds.hist_off)
    # every history length that occurs, plus a spread of users
    users = sorted({int(np.where(hl == h)[0][0]) for h in np.unique(hl)} | set(range(0, ds.n_users, max(1, ds.n_users // 300))))
    cat, lens = b.build(users)
    flat = cat.cpu().numpy()
    o = 0
    for u, n in zip(users, lens):
        want = ds.prompt_ids(u)
        assert n == len(want), (u, n, len(want))
        assert flat[o:o + n].tolist() == want, f"user {u} (history {hl[u]})"
        o += n
    assert o == flat.shape[0]


def test_search_on_device_built_prompts():
    from atspeed_b200.constraint import compile_constraint
    from atspeed_b200.engine import DeviceModel, DeviceTrie, ModelSpec, Session
    from atspeed_b200.prompts import DevicePromptBuilder
    ds = dataset("beauty")
    models = {}
    for which in ("target", "correlated"):
        sh, W = stack_weights("ref_bf16", "beauty", which)
        spec = ModelSpec(sh.vocab, sh.hidden, sh.n_layers, sh.n_heads, sh.head_dim, sh.mlp, sh.eps, sh.rope_theta)
        models[which] = DeviceModel(spec, W, "cuda")
    csr = compile_constraint(constraint_fn("beauty", "strict"), ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
    sess = Session(models["target"], models["correlated"], DeviceTrie(csr, torch.device("cuda")), 10, 40, 4, max_users=8)
    users = [0, 5, 11, 300, 1200, 2500]
    cat, lens = DevicePromptBuilder(ds, "cuda").build(users)
    tok = torch.zeros(len(users), 10, 6, dtype=torch.int32, device="cuda")
    sc = torch.zeros(len(users), 10, dtype=torch.float32, device="cuda")
    sess.bssd_batch_device(cat, lens, 3, tok, sc)
    want = sess.bssd_batch([ds.prompt_ids(u) for u in users], 3)
    got = tok.cpu().numpy()[:, :, :4]
    for i, w in enumerate(want):
        assert got[i].tolist() == w["tokens"].tolist()
