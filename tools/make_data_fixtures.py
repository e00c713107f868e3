"""Derive the compact, travel-safe data fixtures under data/ from the reference's data files.

Reads  /root/reference/data/<ds>/{training,validation,testing}_dict.npy and <ds>.LCRec-1e-3lr.json
(the inputs of SeqRecDataset, reference code/data.py:138-143,232-263) and writes data/<ds>.npz with
  item_codes  int16 [n_items, 4]   numeric suffix of '<a_x>','<b_y>','<c_z>','<d_w>' per item id
  hist_off    int32 [n_users+1]    CSR offsets into hist_items
  hist_items  int32 [...]          last `max_his_len`=20 of train+valid items, per test user
  gt_off/gt_items                  ground-truth test items per test user
  uid         int32 [n_users]      original user ids (dict order, users with a non-empty test list)
Only runs in the build container (the reference is absent on the GPU box); outputs are committed.
"""
import json, os, re, sys
import numpy as np

REF = "/root/reference/data"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "data")
MAX_HIS = 20  # reference code/utils.py:48 (--max_his_len default)


def main():
    for ds in ("beauty", "games"):
        d = os.path.join(REF, ds)
        tr = np.load(os.path.join(d, "training_dict.npy"), allow_pickle=True).item()
        va = np.load(os.path.join(d, "validation_dict.npy"), allow_pickle=True).item()
        te = np.load(os.path.join(d, "testing_dict.npy"), allow_pickle=True).item()
        idx = json.load(open(os.path.join(d, f"{ds}.LCRec-1e-3lr.json")))
        n_items = max(int(k) for k in idx) + 1
        codes = np.full((n_items, 4), -1, np.int16)
        for k, v in idx.items():
            assert len(v) == 4
            for j, (tok, letter) in enumerate(zip(v, "abcd")):
                m = re.fullmatch(rf"<{letter}_(\d+)>", tok)
                assert m, tok
                codes[int(k), j] = int(m.group(1))
        assert (codes >= 0).all()
        uid, hoff, hist, goff, gt = [], [0], [], [0], []
        for u in te:  # dict order == reference _process_test_data order (code/data.py:235)
            if len(te[u]):
                h = list(tr[u]) + list(va[u])
                h = h[-MAX_HIS:]
                uid.append(u)
                hist += h
                hoff.append(len(hist))
                gt += list(te[u])
                goff.append(len(gt))
        np.savez_compressed(os.path.join(OUT, f"{ds}.npz"), item_codes=codes,
                            uid=np.asarray(uid, np.int32), hist_off=np.asarray(hoff, np.int32),
                            hist_items=np.asarray(hist, np.int32), gt_off=np.asarray(goff, np.int32),
                            gt_items=np.asarray(gt, np.int32))
        print(ds, "items", n_items, "users", len(uid), "hist mean", len(hist) / len(uid))


if __name__ == "__main__":
    sys.exit(main())
