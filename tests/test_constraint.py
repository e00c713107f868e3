"""CPU tests of the constraint contract: the Trie mirror behaves like the reference's, and compiling a
`prefix_allowed_tokens_fn` to the CSR child table reproduces the callable's masks bit-exactly."""
import numpy as np
import pytest
import torch

from _common import constraint_fn, dataset
from atspeed_b200.constraint import compile_constraint
from atspeed_b200.generation_trie import Trie, prefix_allowed_tokens_fn

# SURVEY 8a-9 known-answer facts: nodes per depth of the strict item tries
NODES = {"beauty": [1, 91, 6539, 11172, 12023], "games": [1, 248, 11317, 16628, 17289]}


def test_trie_mirror_semantics():
    t = Trie([[1, 5, 6, 2], [1, 5, 7, 2], [1, 8, 9, 2]])
    assert t.get([]) == [1] and t.get([1]) == [5, 8] and t.get([1, 5]) == [6, 7]      # insertion order
    assert t.get([1, 9]) == [] and t[[1, 8, 9]] == [2] and len(t) == 3
    assert sorted(map(tuple, t)) == [(1, 5, 6, 2), (1, 5, 7, 2), (1, 8, 9, 2)]
    t2 = Trie.load_from_dict(t.trie_dict)
    assert len(t2) == 3
    fn = prefix_allowed_tokens_fn(t)
    assert fn(0, torch.tensor([1, 5])) == [6, 7]
    other = Trie([[3, 4]])
    t.append(other, 1)                       # chaining (reference :19-21,55-57,67-68)
    assert t.get([]) == [3] and t.get([3]) == [4]


@pytest.mark.parametrize("name", ["beauty", "games"])
def test_strict_trie_compiles_to_known_node_counts(name):
    ds = dataset(name)
    fn = constraint_fn(name, "strict")
    csr = compile_constraint(fn, ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(7), use_cache=False)
    assert csr.kind == "trie"
    assert csr.n_nodes == sum(NODES[name]) and csr.n_edges == sum(NODES[name][1:])
    assert csr.max_fanout == NODES[name][1] or csr.max_fanout <= 256
    rng = np.random.default_rng(0)
    seqs = ds.item_sequences()
    for row in seqs[rng.integers(0, len(seqs), 200)]:
        node = 0
        for d in range(4):
            prefix = [int(t) for t in row[:d]]
            want = sorted(fn(0, torch.tensor(ds.prompt_ids(3) + prefix)))
            assert csr.children(node).tolist() == want
            node = csr.walk(prefix + [int(row[d])])
            assert node >= 0
    assert csr.walk([int(seqs[0][0]), 31999]) == -1


@pytest.mark.parametrize("name", ["beauty", "games"])
def test_positional_fn_compiles_to_one_node_per_depth(name):
    ds = dataset(name)
    fn = constraint_fn(name, "positional")
    csr = compile_constraint(fn, ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(5), use_cache=False)
    assert csr.kind == "positional" and csr.n_nodes == 5
    for d in range(4):
        assert csr.children(d).tolist() == sorted(ds.positional_allowed()[d])


def test_opaque_callables_are_probed():
    ds = dataset("beauty")
    strict, pos = constraint_fn("beauty", "strict"), constraint_fn("beauty", "positional")
    class Opaque:                                 # hides the dict / trie: forces probing
        def __init__(self, f):
            self._f = f

        def __call__(self, b, s):
            return self._f(b, s)

    opaque_pos = Opaque(pos)
    csr = compile_constraint(opaque_pos, ds.prompt_ids(0), 4, use_cache=False)
    assert csr.kind == "probed-positional"
    ref = compile_constraint(pos, ds.prompt_ids(0), 4, use_cache=False)
    for d in range(4):
        assert csr.children(csr.walk([int(ref.children(k)[0]) for k in range(d)])).tolist() == ref.children(d).tolist()
    # a small real trie behind an opaque callable: probed node by node
    small = Trie([[1, a, b, 2] for a in (32000, 32001, 32005) for b in (32100 + a % 3, 32107)])
    from atspeed_b200.generation_trie import suffix_prefix_allowed_tokens_fn
    from atspeed_b200.prompts import RESPONSE_SEP
    f = suffix_prefix_allowed_tokens_fn(small, RESPONSE_SEP)
    g = Opaque(f)
    c1 = compile_constraint(g, ds.prompt_ids(0), 2, use_cache=False)
    c2 = compile_constraint(f, ds.prompt_ids(0), 2, use_cache=False)
    assert c1.kind == "probed" and c2.kind == "trie"
    np.testing.assert_array_equal(c1.child_tok, c2.child_tok)
    np.testing.assert_array_equal(c1.child_off, c2.child_off)


def test_prompt_dependent_constraint_is_rejected():
    ds = dataset("beauty")
    p0 = ds.prompt_ids(0)
    fn = lambda b, s: [32000 + int(s[40]) % 7, 32050]   # noqa: E731  (depends on a history token of the prompt)
    other = next(ds.prompt_ids(u) for u in range(1, 50) if ds.prompt_ids(u)[40] % 7 != p0[40] % 7)
    with pytest.raises(ValueError):
        compile_constraint(fn, p0, 2, other_prompt=other, use_cache=False)


def test_unconstrained_table_is_cached_per_vocab_and_depth():
    """beamSD.get_session keys device tries and sessions on the table's identity: `prefix_allowed_tokens_fn=None` must not
    produce a fresh table (hence a fresh multi-GB session) per search."""
    from atspeed_b200.constraint import compile_constraint
    a = compile_constraint(None, [1, 2, 3], 2, vocab_size=300)
    b = compile_constraint(None, [9], 2, vocab_size=300)
    assert a is b and a.n_edges == 600
    assert compile_constraint(None, [9], 3, vocab_size=300) is not a
    assert compile_constraint(None, [9], 2, vocab_size=301) is not a
    assert compile_constraint(None, [9], 2, vocab_size=300, use_cache=False) is not a
