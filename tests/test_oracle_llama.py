"""Pins oracle/llama_ref.py (the CPU restatement of the forward the reference reaches through `model(**inputs)`,
code/beamSD.py:52,221) against the installed transformers `LlamaForCausalLM` on the same random weights, through exactly
the call the reference makes: 4-D additive mask (0 = visible, finfo.min = hidden), explicit position_ids, legacy-style KV.
Three calls, as in a BSSD round: the prompt; a beam-tree step on top of the cached prompt (beams on the sequence axis, all at
one depth-based position); a second tree level whose beams see only their own ancestors."""
import pytest
import torch

from oracle import llama_ref as LR

transformers = pytest.importorskip("transformers")


def _hf_model(shape, W):
    from transformers import LlamaConfig, LlamaForCausalLM
    cfg = LlamaConfig(vocab_size=shape.vocab, hidden_size=shape.hidden, intermediate_size=shape.mlp,
                      num_hidden_layers=shape.n_layers, num_attention_heads=shape.n_heads, num_key_value_heads=shape.n_heads,
                      rms_norm_eps=shape.eps, tie_word_embeddings=False, attention_bias=False, mlp_bias=False,
                      max_position_embeddings=512, pad_token_id=0, bos_token_id=1, eos_token_id=2)
    cfg._attn_implementation = "eager"
    m = LlamaForCausalLM(cfg).eval()
    missing = m.load_state_dict(LR.weights_to_hf_state_dict(W), strict=False)
    assert not [k for k in missing.missing_keys if "rotary" not in k], missing
    return m


def _additive(vis):
    return torch.where(vis, 0.0, torch.finfo(torch.float32).min)[None, None].to(torch.float32)


@pytest.mark.parametrize("shape_fn,seed", [(LR.shape_tiny_target, 0), (LR.shape_tiny_draft, 1), (LR.shape_small_draft, 2)])
def test_forward_matches_transformers_prompt_tree_and_cached_calls(shape_fn, seed):
    from transformers.cache_utils import DynamicCache
    V = 211
    shape = shape_fn(V)
    W = LR.make_weights(shape, seed, std=0.08)
    ref = LR.RefLlama(shape, W, "fp32")
    hf = _hf_model(shape, W)
    g = torch.Generator().manual_seed(seed)
    P, n1, n2 = 13, 5, 7
    prompt = torch.randint(3, V, (P,), generator=g)
    lvl1 = torch.randint(3, V, (n1,), generator=g)
    lvl2 = torch.randint(3, V, (n2,), generator=g)
    parent2 = torch.randint(0, n1, (n2,), generator=g)

    # call 1: the prompt, causal
    vis0 = torch.tril(torch.ones(P, P, dtype=torch.bool))
    pos0 = torch.arange(P)
    cache = ref.new_cache()
    mine0 = ref.forward(prompt, pos0, vis0, cache)
    with torch.no_grad():
        out0 = hf(input_ids=prompt[None], attention_mask=_additive(vis0), position_ids=pos0[None], use_cache=True)
    torch.testing.assert_close(mine0, out0.logits[0], atol=2e-5, rtol=1e-5)

    # call 2: n1 sibling beams after the prompt, each sees the prompt and itself; all at position P (code/beamSD.py:87-91)
    vis1 = torch.cat((torch.ones(n1, P, dtype=torch.bool), torch.eye(n1, dtype=torch.bool)), 1)
    pos1 = torch.full((n1,), P)
    mine1 = ref.forward(lvl1, pos1, vis1, cache)
    with torch.no_grad():
        out1 = hf(input_ids=lvl1[None], attention_mask=_additive(vis1), position_ids=pos1[None],
                  past_key_values=out0.past_key_values, use_cache=True)
    torch.testing.assert_close(mine1, out1.logits[0], atol=2e-5, rtol=1e-5)

    # call 3: n2 children, each sees the prompt, its parent among the n1 siblings, and itself; position P + 1.  The cache
    # is handed over the way the reference does after verify: rebuilt as a list of (k, v) (code/beamSD.py:418-429)
    vis2 = torch.zeros(n2, P + n1 + n2, dtype=torch.bool)
    vis2[:, :P] = True
    vis2[torch.arange(n2), P + parent2] = True
    vis2[torch.arange(n2), P + n1 + torch.arange(n2)] = True
    pos2 = torch.full((n2,), P + 1)
    mine2 = ref.forward(lvl2, pos2, vis2, cache)
    legacy = [(k, v) for k, v, *_ in out1.past_key_values]
    with torch.no_grad():
        out2 = hf(input_ids=lvl2[None], attention_mask=_additive(vis2), position_ids=pos2[None],
                  past_key_values=DynamicCache(ddp_cache_data=legacy), use_cache=True)
    torch.testing.assert_close(mine2, out2.logits[0], atol=2e-5, rtol=1e-5)
    assert len(cache) == P + n1 + n2
    # the oracle's cache is the reference's KV in [S, H, D] layout
    k_hf = list(out2.past_key_values)[0][0][0].permute(1, 0, 2)
    torch.testing.assert_close(cache.k[0], k_hf, atol=2e-5, rtol=1e-5)


def test_logit_rows_selects_rows_after_the_final_norm():
    shape = LR.shape_tiny_target(97)
    ref = LR.RefLlama(shape, LR.make_weights(shape, 3, std=0.08), "fp32")
    toks, pos = torch.arange(3, 12), torch.arange(9)
    vis = torch.tril(torch.ones(9, 9, dtype=torch.bool))
    full = ref.forward(toks, pos, vis, ref.new_cache())
    rows = torch.tensor([8, 2])
    torch.testing.assert_close(ref.forward(toks, pos, vis, ref.new_cache(), logit_rows=rows), full[rows])


def test_bf16_contract_is_close_to_an_hf_bf16_module():
    """precision="bf16" rounds where an HF bf16 module rounds.  Against the real bf16 module on CPU (whose GEMMs accumulate
    in their own order) the logits agree to a couple of bf16 ulps: measured max |diff| 0.0625 = 2 ulp at |logit| ~ 4.8, 22 % of
    the logits bit-equal.  Tolerance here: 3 % of the largest logit."""
    V = 157
    shape = LR.shape_tiny_target(V)
    W = LR.make_weights(shape, 4, std=0.16, dtype=torch.bfloat16)
    hf = _hf_model(shape, W).to(torch.bfloat16)
    P = 11
    prompt = torch.randint(3, V, (P,), generator=torch.Generator().manual_seed(9))
    vis = torch.tril(torch.ones(P, P, dtype=torch.bool))
    pos = torch.arange(P)
    mine = LR.RefLlama(shape, W, "bf16", round_logits=True).forward(prompt, pos, vis, LR.RefCache())
    with torch.no_grad():
        out = hf(input_ids=prompt[None], attention_mask=_additive(vis).to(torch.bfloat16), position_ids=pos[None])
    theirs = out.logits[0].float()
    scale = theirs.abs().max().item()
    assert (mine - theirs).abs().max().item() <= 0.03 * max(1.0, scale), ((mine - theirs).abs().max().item(), scale)
