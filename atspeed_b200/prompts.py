"""Item-code vocabulary, test users and prompts for the Beauty / Games fixtures.

What the reference does (file:line relative to /root/reference):
  * code/finetune_llama.py:84 + code/data.py:46-57 -- the distinct code strings ('<a_12>', ...) are
    sorted lexicographically and appended to the 32000-entry LLaMA vocabulary, so a code token's id
    is 32000 + its rank in that sorted list (Beauty: 859 tokens, V=32859; Games: 1014, V=33014).
  * code/data.py:232-263 -- a test prompt is the SFT template around the user's last 20 history
    items (train+valid), every item written as its 4 code tokens, items separated by ", ".
  * code/inference.py:130 -- items are keyed in the strict trie as [BOS, a, b, c, d, EOS].

The LLaMA sentencepiece model is not available offline, so the *text* part of the prompt is mapped
to fixed synthetic ids in [3, 31999]; code-token ids are exact.  The prompt ends with the id run
`RESPONSE_SEP` standing for the tokens of "Response:", which both constraint functions search for.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from typing import Dict, List, Sequence

import numpy as np

PAD_ID, BOS_ID, EOS_ID = 0, 1, 2  # reference code/inference.py:105-107
BASE_VOCAB = 32000
RESPONSE_SEP = (13291, 29901)  # stands for tokenizer("Response:")["input_ids"][1:]
MAX_HIS_LEN = 20  # reference code/utils.py:48

_DATA_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")


def _text_ids(n: int, seed: int) -> List[int]:
    """Deterministic stand-in ids for template text (never a code token, BOS/EOS/PAD or the SEP run)."""
    out, x = [], seed
    while len(out) < n:
        x = (x * 1103515245 + 12345) & 0x7FFFFFFF
        t = 3 + x % (BASE_VOCAB - 3)
        if t not in RESPONSE_SEP:
            out.append(t)
    return out


_PREFIX = _text_ids(38, 7)     # "Below is an instruction ... ### Instruction:\nThe user has interacted with items"
_SUFFIX = _text_ids(20, 11)    # "in chronological order. Can you predict ... may expect?\n\n###"
_SEP1 = _text_ids(1, 13)       # ","
_SEP2 = _text_ids(2, 17)       # ", " tokenised as two pieces on alternating items


@dataclass
class RecDataset:
    """Derived fixture of one reference dataset (see tools/make_data_fixtures.py)."""
    name: str
    item_codes: np.ndarray      # [n_items, 4] int16: numeric suffix per level
    uid: np.ndarray
    hist_off: np.ndarray
    hist_items: np.ndarray
    gt_off: np.ndarray
    gt_items: np.ndarray
    level_tokens: List[List[str]]          # per level, sorted distinct code strings of that level
    token_id: Dict[str, int]               # code string -> vocabulary id
    item_token_ids: np.ndarray             # [n_items, 4] int32 vocabulary ids

    @property
    def vocab_size(self) -> int:
        return BASE_VOCAB + len(self.token_id)

    @property
    def n_users(self) -> int:
        return len(self.uid)

    def level_ranges(self):
        """[(lo, hi)] inclusive id range per level (contiguous because the sort groups by letter)."""
        out = []
        for lv in self.level_tokens:
            ids = [self.token_id[t] for t in lv]
            out.append((min(ids), max(ids)))
        return out

    def history(self, u: int) -> np.ndarray:
        return self.hist_items[self.hist_off[u]:self.hist_off[u + 1]]

    def ground_truth(self, u: int) -> np.ndarray:
        return self.gt_items[self.gt_off[u]:self.gt_off[u + 1]]

    def prompt_ids(self, u: int) -> List[int]:
        """Synthetic tokenisation of the reference test prompt of user index `u` (code/data.py:246-256)."""
        ids = [BOS_ID] + list(_PREFIX)
        h = self.history(u)
        for j, it in enumerate(h):
            ids += [int(t) for t in self.item_token_ids[it]]
            if j + 1 < len(h):
                ids += _SEP1 if j % 2 == 0 else _SEP2
        ids += list(_SUFFIX) + list(RESPONSE_SEP)
        return ids

    def item_sequences(self) -> np.ndarray:
        """Distinct 4-token item code sequences, sorted (12023 for Beauty, 17289 for Games)."""
        return np.unique(self.item_token_ids, axis=0)

    def strict_trie_sequences(self) -> List[List[int]]:
        """[BOS, a, b, c, d, EOS] per distinct item, the keys of the strict trie (code/inference.py:130)."""
        return [[BOS_ID] + [int(t) for t in row] + [EOS_ID] for row in self.item_sequences()]

    def positional_allowed(self) -> Dict[int, List[int]]:
        """{depth: sorted allowed ids} as built by get_prefix_allowed_tokens_fn (code/data.py:86-94)."""
        out = {i: sorted(int(x) for x in np.unique(self.item_token_ids[:, i])) for i in range(4)}
        out[4] = [EOS_ID]
        return out

    def decode_items(self, token_rows: Sequence[Sequence[int]]) -> List[str]:
        """Vocabulary ids of the 4 generated tokens -> the item's code string ('<a_1><b_2><c_3><d_4>')."""
        inv = getattr(self, "_inv", None)
        if inv is None:
            inv = {v: k for k, v in self.token_id.items()}
            object.__setattr__(self, "_inv", inv)
        return ["".join(inv.get(int(t), f"<unk_{int(t)}>") for t in row) for row in token_rows]

    def ground_truth_strings(self, u: int) -> List[str]:
        return self.decode_items(self.item_token_ids[self.ground_truth(u)])


def prompt_len(n_history: int) -> int:
    """Length of the prompt of a user with `n_history` history items (closed form of RecDataset.prompt_ids)."""
    h = n_history
    seps = ((h - 1 + 1) // 2) * len(_SEP1) + ((h - 1) // 2) * len(_SEP2) if h > 0 else 0
    return 1 + len(_PREFIX) + 4 * h + seps + len(_SUFFIX) + len(RESPONSE_SEP)


class DevicePromptBuilder:
    """Prompts built ON THE DEVICE from history item ids (csrc/prompt.cu): the dataset's item -> 4 code-token
    table and the concatenated user histories live in HBM; `build(users)` launches one kernel that writes the users' prompts
    into one concatenated int32 tensor -- the input layout of Session.bssd_batch_device -- and returns it with the prompt
    lengths (host closed form, no tokenisation).  Mirrors reference code/data.py:232-263 + code/collator.py:50-75."""

    def __init__(self, ds: RecDataset, device):
        import ctypes as C

        import torch

        from . import _lib
        self.ds, self.device, self.lib, self._C, self._torch = ds, torch.device(device), _lib.load(), C, torch
        self.item_tok = torch.from_numpy(np.ascontiguousarray(ds.item_token_ids.astype(np.int32))).to(self.device)
        self.hist_items = torch.from_numpy(np.ascontiguousarray(ds.hist_items.astype(np.int32))).to(self.device)
        t = _lib.PromptTemplate()
        t.bos, t.n_prefix, t.n_suffix, t.n_resp = BOS_ID, len(_PREFIX), len(_SUFFIX), len(RESPONSE_SEP)
        t.n_sep_even, t.n_sep_odd = len(_SEP1), len(_SEP2)
        for i, v in enumerate(list(_PREFIX) + list(_SUFFIX) + list(RESPONSE_SEP) + list(_SEP1) + list(_SEP2)):
            t.ids[i] = int(v)
        self.template = t

    def build(self, users: Sequence[int]):
        torch, C = self._torch, self._C
        ds = self.ds
        begin = np.asarray([ds.hist_off[u] for u in users], dtype=np.int64)
        hlen = np.asarray([ds.hist_off[u + 1] - ds.hist_off[u] for u in users], dtype=np.int32)
        lens = [prompt_len(int(h)) for h in hlen]
        off = np.zeros(len(users), dtype=np.int64)
        off[1:] = np.cumsum(lens)[:-1]
        meta = torch.from_numpy(np.concatenate([begin, off])).to(self.device, non_blocking=True)
        hl = torch.from_numpy(hlen).to(self.device, non_blocking=True)
        out = torch.empty(int(sum(lens)), dtype=torch.int32, device=self.device)
        n = len(users)
        from . import _lib
        with torch.cuda.device(self.device):
            _lib.check(self.lib.atspeed_build_prompts(self.item_tok.data_ptr(), self.hist_items.data_ptr(), meta[:n].data_ptr(),
                                                      hl.data_ptr(), meta[n:].data_ptr(), n, C.byref(self.template), out.data_ptr(),
                                                      C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)))
        return out, lens


def load_dataset(name: str, data_dir: str | None = None) -> RecDataset:
    z = np.load(os.path.join(data_dir or _DATA_DIR, f"{name}.npz"))
    codes = z["item_codes"]
    strings = set()
    level_tokens = []
    for j, letter in enumerate("abcd"):
        lv = sorted({f"<{letter}_{int(c)}>" for c in np.unique(codes[:, j])})
        level_tokens.append(lv)
        strings.update(lv)
    ordered = sorted(strings)  # lexicographic, as tokenizer.add_tokens(sorted(new_tokens))
    token_id = {s: BASE_VOCAB + r for r, s in enumerate(ordered)}
    item_token_ids = np.empty(codes.shape, np.int32)
    for j, letter in enumerate("abcd"):
        lut = {int(c): token_id[f"<{letter}_{int(c)}>"] for c in np.unique(codes[:, j])}
        item_token_ids[:, j] = [lut[int(c)] for c in codes[:, j]]
    return RecDataset(name, codes, z["uid"], z["hist_off"], z["hist_items"], z["gt_off"], z["gt_items"],
                      level_tokens, token_id, item_token_ids)
