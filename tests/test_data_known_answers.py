"""Known-answer facts of the data layer that the hot path depends on (SURVEY.md sections 8a-0, 8a-9, 8a-9', 8c): the code-token
id layout `tokenizer.add_tokens(sorted(new_tokens))` gives (/root/reference/code/finetune_llama.py:84, code/data.py:46-57),
the number of test users, and the shape of the strict item trie -- all derived by the survey from the reference's own data
files; data/*.npz must reproduce them."""
import numpy as np
import pytest

from _common import constraint_fn, dataset

FACTS = {
    "beauty": dict(users=3553, V=32859, ranges=[(32000, 32090), (32091, 32346), (32347, 32602), (32603, 32858)],
                   sequences=12023, nodes_per_depth=[1, 91, 6539, 11172, 12023], edges_with_eos=41848,
                   positional=[91, 256, 256, 256]),
    "games": dict(users=8696, V=33014, ranges=[(32000, 32247), (32248, 32503), (32504, 32757), (32758, 33013)],
                  sequences=17289, nodes_per_depth=[1, 248, 11317, 16628, 17289], edges_with_eos=62771,
                  positional=[248, 256, 254, 256]),
}


@pytest.mark.parametrize("name", ["beauty", "games"])
def test_token_layout_users_and_trie_shape(name):
    from atspeed_b200.constraint import compile_constraint
    ds, f = dataset(name), FACTS[name]
    assert ds.n_users == f["users"] and ds.vocab_size == f["V"]
    assert [tuple(r) for r in ds.level_ranges()] == f["ranges"]
    seqs = ds.strict_trie_sequences()
    assert len(seqs) == f["sequences"]
    # items are keyed [BOS, a, b, c, d, EOS] (code/inference.py:130), one code token per level range
    for s in seqs[:200]:
        assert len(s) == 6 and s[0] == 1 and s[-1] == 2
        assert all(lo <= t <= hi for t, (lo, hi) in zip(s[1:5], f["ranges"]))
    fn = constraint_fn(name, "strict")
    csr = compile_constraint(fn, ds.prompt_ids(0), 4, other_prompt=ds.prompt_ids(1))
    depth = np.zeros(csr.n_nodes, dtype=np.int64)
    for node in range(csr.n_nodes):
        for e in range(csr.child_off[node], csr.child_off[node + 1]):
            if csr.child_node[e] >= 0:
                depth[csr.child_node[e]] = depth[node] + 1
    assert np.bincount(depth).tolist() == f["nodes_per_depth"] and csr.n_nodes == sum(f["nodes_per_depth"])
    # 4 generated tokens: every non-root node has one incoming edge; the EOS edges of the 5th step make up the survey's total
    assert csr.n_edges == csr.n_nodes - 1 and csr.n_edges + f["sequences"] == f["edges_with_eos"]
    # the positional constraint (code/data.py:84-104): allowed set depends on depth only, then {EOS}
    pos = ds.positional_allowed()
    assert [len(pos[d]) for d in range(4)] == f["positional"] and list(pos[4]) == [2]
    for d, (lo, hi) in enumerate(f["ranges"]):
        assert min(pos[d]) >= lo and max(pos[d]) <= hi


def test_prompt_length_closed_form_matches_the_tokenisation():
    """The device-side prompt builder (csrc/prompt.cu) gets its prompt lengths from prompts.prompt_len: it must equal the
    length of the host tokenisation for every test user of both datasets."""
    import numpy as np
    from atspeed_b200.prompts import load_dataset, prompt_len
    for name in ("beauty", "games"):
        ds = load_dataset(name)
        hl = np.diff(ds.hist_off)
        for u in range(0, ds.n_users, 7):
            assert prompt_len(int(hl[u])) == len(ds.prompt_ids(u)), (name, u)
