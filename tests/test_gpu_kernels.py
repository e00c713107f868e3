"""GPU parity tests of the stand-alone kernels, all called through the C ABI (ctypes).

Kernel (a) and (c) are integer/index work and must be bit-exact against the oracle (scores of (a) within
1e-5 absolute: the log-sum-exp is summed in a different order than torch's).  The tcgen05 GEMM and the
tree attention are floating point: compared with a plain torch fp32 reference of the same op on the same
bf16-valued inputs, tolerance stated per test.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def lib():
    from atspeed_b200 import _lib
    return _lib.load()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _check(lib, rc):
    assert rc == 0, lib.atspeed_last_error().decode()


# ---------------------------------------------------------------------------------------------------
# tcgen05 GEMM
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,K,rows", [(1, 64, (128,)), (10, 4096, (4096, 4096, 4096)), (16, 768, (768, 768, 768)),
                                      (37, 256, (256, 256, 256)), (50, 4096, (11008, 11008)), (121, 4096, (32859,)),
                                      (130, 11008, (4096,)), (220, 4096, (4096,)), (220, 4096, (11008, 11008)),
                                      (256, 128, (64, 64, 64)), (289, 4096, (4096, 4096, 4096)), (300, 3072, (768,)),
                                      (512, 64, (40,)), (90, 768, (3072, 3072)), (7, 192, (200, 72))])
def test_gemm_tcgen05_matches_torch(lib, T, K, rows):
    """tcgen05 GEMM with persistent CTAs, K-cut tiles and fixed-order slice reduction vs torch fp32."""
    g = torch.Generator(device="cuda").manual_seed(T * 1000 + K)
    x = (torch.randn(T, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    ws = [(torch.randn(r, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16) for r in rows]
    ldo = sum(rows) + 3
    out = torch.full((T, ldo), float("nan"), device="cuda", dtype=torch.float32)
    ptr = [w.data_ptr() for w in ws] + [None] * (3 - len(ws))
    rr = list(rows) + [0] * (3 - len(rows))
    nbytes = C.c_size_t(0)
    _check(lib, lib.atspeed_gemm_scratch_bytes(T, K, rr[0], rr[1], rr[2], C.byref(nbytes)))
    scratch = torch.full((nbytes.value // 4,), float("nan"), device="cuda", dtype=torch.float32)
    outs = []
    for _ in range(2):
        _check(lib, lib.atspeed_gemm_bf16(x.data_ptr(), T, K, ptr[0], rr[0], ptr[1], rr[1], ptr[2], rr[2],
                                          scratch.data_ptr(), out.data_ptr(), ldo, _stream()))
        torch.cuda.synchronize()
        outs.append(out.clone())
    got = out[:, : sum(rows)]
    ref = x.float() @ torch.cat(ws).float().T
    # fp32 accumulation of exact bf16 products: only the summation order differs
    err = (got - ref).abs().max().item()
    assert torch.isfinite(got).all(), "unwritten outputs"
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    assert torch.isnan(out[:, sum(rows):]).all(), "wrote past the output columns"
    assert torch.equal(outs[0][:, : sum(rows)], outs[1][:, : sum(rows)]), "run-to-run nondeterminism"


@pytest.mark.parametrize("T,K,rows", [(289, 4096, (4096, 4096, 4096)), (300, 3072, (768,)), (400, 4096, (11008, 11008)),
                                      (481, 4096, (32859,)), (512, 768, (3072, 3072)), (512, 11008, (4096,))])
def test_gemm_cluster_of_four_matches_torch(lib, T, K, rows, monkeypatch):
    """ATSPEED_GEMM_CLUSTER=4 (opt-in): two CTA pairs per cluster on adjacent tiles, the activation tiles fetched once per cluster
    and multicast to the twin CTA.  Every output element is the same sequence of MMAs over the same k-blocks as in the pair
    kernel; only the cut points of the k-range differ (fewer workers), so the result is compared with torch fp32 and must be
    deterministic run to run.  (On a B200 33 clusters of 4 can be resident; the plan uses that many workers.)"""
    monkeypatch.setenv("ATSPEED_GEMM_CLUSTER", "4")
    g = torch.Generator(device="cuda").manual_seed(T * 7 + K)
    x = (torch.randn(T, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16)
    ws = [(torch.randn(r, K, generator=g, device="cuda") * 0.5).to(torch.bfloat16) for r in rows]
    cols = sum(rows)
    ptr = [w.data_ptr() for w in ws] + [None] * (3 - len(ws))
    rr = list(rows) + [0] * (3 - len(rows))
    nbytes = C.c_size_t(0)
    _check(lib, lib.atspeed_gemm_scratch_bytes(T, K, rr[0], rr[1], rr[2], C.byref(nbytes)))
    scratch = torch.full((nbytes.value // 4,), float("nan"), device="cuda", dtype=torch.float32)
    outs = []
    for _ in range(2):
        out = torch.full((T, cols), float("nan"), device="cuda", dtype=torch.float32)
        _check(lib, lib.atspeed_gemm_bf16(x.data_ptr(), T, K, ptr[0], rr[0], ptr[1], rr[1], ptr[2], rr[2],
                                          scratch.data_ptr(), out.data_ptr(), cols, _stream()))
        torch.cuda.synchronize()
        outs.append(out)
    ref = x.float() @ torch.cat(ws).float().T
    assert torch.isfinite(outs[0]).all(), "unwritten outputs"
    err = (outs[0] - ref).abs().max().item()
    assert err <= 2e-3 * max(1.0, ref.abs().max().item()), f"max abs err {err}"
    assert torch.equal(outs[0], outs[1]), "run-to-run nondeterminism"


# ---------------------------------------------------------------------------------------------------
# kernel (a)
# ---------------------------------------------------------------------------------------------------
def _device_trie(csr):
    from atspeed_b200.engine import DeviceTrie
    return DeviceTrie(csr, torch.device("cuda"))


def _oracle_topk(logits_f32, node, csr, B):
    """CPU oracle for one row: log_softmax over the full vocabulary, keep the node's children, rank by
    (logp desc, token asc), drop non-finite."""
    lp = torch.log_softmax(logits_f32.double(), -1)
    toks = csr.children(node)
    vals = lp[torch.from_numpy(toks.astype(np.int64))]
    order = sorted(range(len(toks)), key=lambda i: (-float(vals[i]), int(toks[i])))
    order = [i for i in order if np.isfinite(float(vals[i]))][:B]
    return [int(toks[i]) for i in order], [float(vals[i]) for i in order]


@pytest.mark.parametrize("dtype", ["fp32", "bf16"])
@pytest.mark.parametrize("ds_name,kind", [("beauty", "strict"), ("games", "positional")])
def test_mask_logsoftmax_topk(lib, dtype, ds_name, kind):
    from atspeed_b200.constraint import compile_constraint
    ds = dataset(ds_name)
    csr = compile_constraint(constraint_fn(ds_name, kind), ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(3))
    trie = _device_trie(csr)
    V, rows, B = ds.vocab_size, 45, 40
    g = torch.Generator().manual_seed(5)
    logits = torch.randn(rows, V, generator=g) * 2.0
    logits[3, ds.level_ranges()[0][0] + 2] = float("-inf")      # a masked-out allowed token
    if dtype == "bf16":
        logits = logits.to(torch.bfloat16)
    rng = np.random.default_rng(1)
    depth_nodes = [0] + [int(x) for x in rng.integers(0, csr.n_nodes, rows - 1)]
    depth_nodes[3] = 0
    depth_nodes[7] = -1                                          # padding row
    row_node = torch.tensor(depth_nodes, dtype=torch.int32, device="cuda")
    dl = logits.cuda()                                            # row stride = V (odd): exercises the unaligned path
    ct = torch.zeros(rows * B, dtype=torch.int32, device="cuda")
    ce = torch.zeros_like(ct)
    cl = torch.zeros(rows * B, dtype=torch.float32, device="cuda")
    cc = torch.zeros(rows, dtype=torch.int32, device="cuda")
    lse = torch.zeros(rows, dtype=torch.float32, device="cuda")
    _check(lib, lib.atspeed_mask_logsoftmax_topk(dl.data_ptr(), 1 if dtype == "bf16" else 0, rows, V, V, row_node.data_ptr(),
                                                 None, C.byref(trie.desc), B, ct.data_ptr(), ce.data_ptr(), cl.data_ptr(),
                                                 cc.data_ptr(), lse.data_ptr(), _stream()))
    torch.cuda.synchronize()
    ct, ce, cl, cc, lse = (t.cpu().numpy() for t in (ct, ce, cl, cc, lse))
    lf = logits.float()
    for r in range(rows):
        if depth_nodes[r] < 0:
            assert cc[r] == 0
            continue
        toks, vals = _oracle_topk(lf[r], depth_nodes[r], csr, B)
        assert cc[r] == len(toks), (r, cc[r], len(toks))
        got = ct[r * B: r * B + cc[r]].tolist()
        assert got == toks, (r, got[:5], toks[:5])                               # bit-exact selection and order
        np.testing.assert_allclose(cl[r * B: r * B + cc[r]], vals, atol=2e-5, rtol=0)
        assert all(csr.child_tok[e] == t for e, t in zip(ce[r * B: r * B + cc[r]], got))
        np.testing.assert_allclose(lse[r], float(torch.logsumexp(lf[r].double(), -1)), atol=2e-5)


def test_mask_logsoftmax_topk_unconstrained_wide_node(lib):
    from atspeed_b200.constraint import compile_constraint
    V, rows, B = 5000, 3, 10
    csr = compile_constraint(None, [1, 2, 3], 2, vocab_size=V)
    trie = _device_trie(csr)
    logits = torch.randn(rows, V, generator=torch.Generator().manual_seed(0))
    row_node = torch.zeros(rows, dtype=torch.int32, device="cuda")
    dl = logits.cuda()
    ct = torch.zeros(rows * B, dtype=torch.int32, device="cuda")
    ce = torch.zeros_like(ct)
    cl = torch.zeros(rows * B, dtype=torch.float32, device="cuda")
    cc = torch.zeros(rows, dtype=torch.int32, device="cuda")
    _check(lib, lib.atspeed_mask_logsoftmax_topk(dl.data_ptr(), 0, rows, V, V, row_node.data_ptr(), None, C.byref(trie.desc),
                                                 B, ct.data_ptr(), ce.data_ptr(), cl.data_ptr(), cc.data_ptr(), None, _stream()))
    torch.cuda.synchronize()
    for r in range(rows):
        toks, vals = _oracle_topk(logits[r], 0, csr, B)
        assert ct.cpu()[r * B:(r + 1) * B].tolist() == toks
        np.testing.assert_allclose(cl.cpu()[r * B:(r + 1) * B], vals, atol=2e-5)


# ---------------------------------------------------------------------------------------------------
# kernel (c)
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("row_bytes,planes,S", [(128, 4, 64), (8192, 6, 300), (1536, 4, 200), (32768 + 4096, 2, 40)])
def test_kv_gather_matches_index_copy(lib, row_bytes, planes, S):
    g = torch.Generator(device="cuda").manual_seed(row_bytes)
    buf = torch.randint(0, 255, (planes, S, row_bytes), dtype=torch.uint8, device="cuda", generator=g)
    ref = buf.clone()
    n = 17
    perm = torch.randperm(S)
    src = perm[:n].to(torch.int32)
    dst = perm[n:2 * n].to(torch.int32)          # disjoint from src
    ref[:, dst.long()] = ref[:, src.long()]
    ds_, dd_ = src.cuda(), dst.cuda()
    n_dev = torch.tensor([n], dtype=torch.int32, device="cuda")
    _check(lib, lib.atspeed_kv_gather(buf.data_ptr(), buf.data_ptr(), S * row_bytes, S * row_bytes, planes, row_bytes,
                                      ds_.data_ptr(), dd_.data_ptr(), n_dev.data_ptr(), n + 5, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(buf, ref)
    # out-of-place HF-style beam reorder: dst[b] = src[beam_idx[b]] over whole (position) rows
    out = torch.zeros_like(buf)
    idx = torch.randint(0, S, (S,), generator=torch.Generator().manual_seed(1)).to(torch.int32).cuda()
    ar = torch.arange(S, dtype=torch.int32, device="cuda")
    _check(lib, lib.atspeed_kv_gather(buf.data_ptr(), out.data_ptr(), S * row_bytes, S * row_bytes, planes, row_bytes,
                                      idx.data_ptr(), ar.data_ptr(), None, S, _stream()))
    torch.cuda.synchronize()
    assert torch.equal(out, buf[:, idx.long()])


# ---------------------------------------------------------------------------------------------------
# tree attention
# ---------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("T,S,H,D,P", [(20, 60, 4, 16, 30), (130, 300, 2, 64, 120), (70, 200, 3, 128, 100), (5, 40, 2, 32, 0)])
def test_tree_attention_matches_torch(lib, T, S, H, D, P):
    g = torch.Generator(device="cuda").manual_seed(T + S)
    q = torch.randn(T, H * D, generator=g, device="cuda").to(torch.bfloat16)
    k = torch.randn(S, H * D, generator=g, device="cuda").to(torch.bfloat16)
    v = torch.randn(S, H * D, generator=g, device="cuda").to(torch.bfloat16)
    rng = np.random.default_rng(T)
    prefix = rng.integers(0, P + 1, T).astype(np.int32)
    vis_bits = rng.random((T, S - P)) < 0.1
    vis_bits[np.arange(T), rng.integers(0, S - P, T)] = True     # every query sees something
    words = np.zeros((T, 16), dtype=np.uint32)
    for t in range(T):
        for j in np.nonzero(vis_bits[t])[0]:
            words[t, j >> 5] |= np.uint32(1) << np.uint32(j & 31)
    mask = np.zeros((T, S), dtype=bool)
    for t in range(T):
        mask[t, : prefix[t]] = True
        mask[t, P:] |= vis_bits[t]
    out = torch.zeros(T, H * D, dtype=torch.bfloat16, device="cuda")
    pl = torch.from_numpy(prefix).cuda()
    vw = torch.from_numpy(words.view(np.int32)).cuda()
    _check(lib, lib.atspeed_tree_attention(q.data_ptr(), k.data_ptr(), v.data_ptr(), pl.data_ptr(), vw.data_ptr(), P, T, S,
                                           H, D, out.data_ptr(), _stream()))
    torch.cuda.synchronize()
    qf, kf, vf = (x.float().view(-1, H, D) for x in (q, k, v))
    sc = torch.einsum("thd,shd->hts", qf, kf) / D ** 0.5
    sc = sc.masked_fill(~torch.from_numpy(mask).cuda()[None], float("-inf"))
    ref = torch.einsum("hts,shd->thd", torch.softmax(sc, -1), vf).reshape(T, H * D)
    # output is bf16: allow one bf16 ulp (2^-8 relative) plus fp32 summation noise
    err = (out.float() - ref).abs()
    tol = 1e-2 * ref.abs() + 2e-3
    assert (err <= tol).all(), f"max err {err.max().item()} at {int(err.argmax())}"
