"""Prefix trie over token-id sequences and the HF `prefix_allowed_tokens_fn` contract.

Host-side mirror of /root/reference/code/generation_trie.py (same names, argument meaning and
return values), so code written against the reference keeps working:

  * `Trie(sequences)`, `.trie_dict` (nested dict keyed by token id), `.add`, `.get(prefix)` ->
    children of the node reached by walking `prefix` from the root, in insertion order, `[]` when the
    prefix is absent (reference :47-70); `.append(trie, bos_token_id)` chaining (:19-21,55-57,67-68);
    `load_from_dict`, iteration over stored sequences, `len`, `trie[prefix]`.
  * `prefix_allowed_tokens_fn(trie)` -> `(batch_id, sentence) -> List[int]` walking the WHOLE
    sentence (reference :92-98).
  * `suffix_prefix_allowed_tokens_fn(trie, sep, bos)` -- the working way to key the strict item trie
    (reference code/generate_teacher_data.py:174-188): find the last "Response:" id run, walk
    `[bos] + generated suffix`.

The walk is iterative (no recursion / list slicing per level).  On the device the same trie is a CSR
child table, see atspeed_b200/constraint.py.
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Optional, Sequence


class Trie(object):
    def __init__(self, sequences: Optional[Iterable[Sequence[int]]] = None):
        self.trie_dict: Dict[int, dict] = {}
        self.len = 0
        self.append_trie: Optional["Trie"] = None
        self.bos_token_id: Optional[int] = None
        for seq in sequences or ():
            self.add(seq)

    def append(self, trie: "Trie", bos_token_id: int) -> None:
        self.append_trie, self.bos_token_id = trie, bos_token_id

    def add(self, sequence: Sequence[int]) -> None:
        node = self.trie_dict
        for tok in sequence:
            node = node.setdefault(tok, {})
        self.len += 1

    def get(self, prefix_sequence: Sequence[int]) -> List[int]:
        node = self.trie_dict
        for depth, tok in enumerate(prefix_sequence):
            nxt = node.get(tok)
            if nxt is None:
                if self.append_trie is not None:
                    return self.append_trie.get(list(prefix_sequence)[depth:])
                return []
            node = nxt
        out = list(node.keys())
        if self.append_trie is not None and self.bos_token_id in out:
            out.remove(self.bos_token_id)
            out += list(self.append_trie.trie_dict.keys())
        return out

    @staticmethod
    def load_from_dict(trie_dict: Dict[int, dict]) -> "Trie":
        trie = Trie()
        trie.trie_dict = trie_dict
        trie.len = sum(1 for _ in trie)
        return trie

    def __iter__(self) -> Iterator[List[int]]:
        stack = [([], self.trie_dict)]
        while stack:
            prefix, node = stack.pop()
            if not node:
                yield prefix
                continue
            for tok in reversed(list(node.keys())):
                stack.append((prefix + [tok], node[tok]))

    def __len__(self) -> int:
        return self.len

    def __getitem__(self, value: Sequence[int]) -> List[int]:
        return self.get(value)


def prefix_allowed_tokens_fn(candidate_trie: Trie):
    def prefix_allowed_tokens(batch_id, sentence):
        return candidate_trie.get(sentence.tolist())

    prefix_allowed_tokens.candidate_trie = candidate_trie
    return prefix_allowed_tokens


def suffix_prefix_allowed_tokens_fn(candidate_trie: Trie, sep: Sequence[int], bos_token_id: int = 1):
    sep = list(sep)

    def prefix_allowed_tokens(batch_id, sentence):
        s = sentence.tolist()
        n, m = len(s), len(sep)
        for i in range(n, m - 1, -1):
            if s[i - m:i] == sep:
                return candidate_trie.get([bos_token_id] + s[i:])
        return []

    prefix_allowed_tokens.candidate_trie = candidate_trie
    prefix_allowed_tokens.trie_root_prefix = [bos_token_id]
    return prefix_allowed_tokens


def positional_prefix_allowed_tokens_fn(allowed_tokens: Dict[int, Iterable[int]], sep: Sequence[int]):
    """The constraint inference.py really passes (reference code/data.py:84-104): the allowed set
    depends only on how many tokens follow the last "Response:" run."""
    sep_rev = list(sep)[::-1]
    allowed = {d: list(v) for d, v in allowed_tokens.items()}

    def prefix_allowed_tokens(batch_id, sentence):
        rev = sentence.tolist()[::-1]
        m = len(sep_rev)
        for i in range(len(rev)):
            if rev[i:i + m] == sep_rev:
                return list(allowed[i])
        return None

    prefix_allowed_tokens.allowed_tokens = allowed
    return prefix_allowed_tokens
