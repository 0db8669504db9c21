"""Freeze the reference's own training / test samples for a small synthetic data set (build container only).

    python -m oracle.make_dataset_golden      # writes tests/golden/dataset_smb.json

TEST INFRASTRUCTURE ONLY.  Writes a data set in the reference's file format (gamer_b200.dataset.write_synthetic_files,
fixed seed), runs the UNMODIFIED `SMBExplicitDatasetForDecoder(augment=4)` of SeqRec/datasets/SMB_dataset.py on it in
train and test mode, and stores every sample as token strings + the id lists the class emits.  The test rebuilds the
same files, loads them with gamer_b200.dataset and must reproduce these samples through PackedSessions + collate.
"""
from __future__ import annotations

import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle import ref_shim  # noqa: E402
from gamer_b200 import dataset as ds  # noqa: E402

SPEC = dict(n_users=24, n_items=300, seed=3)
MAX_HIS_LEN = 12


def main():
    ref_shim.load_reference()
    from SeqRec.datasets.SMB_dataset import SMBExplicitDatasetForDecoder
    out = {"spec": SPEC, "max_his_len": MAX_HIS_LEN, "augment": 4}
    with tempfile.TemporaryDirectory() as tmp:
        ds.write_synthetic_files(tmp, "toy", **SPEC)
        for mode in ("train", "test"):
            d = SMBExplicitDatasetForDecoder(augment=4 if mode == "train" else None, dataset="toy", data_path=tmp,
                                             max_his_len=MAX_HIS_LEN, index_file=".index.json", mode=mode)
            out[mode] = [{k: v for k, v in s.items() if k in ("item", "inters", "session_ids", "extended_session_ids",
                                                              "actions", "behavior")} for s in d.inter_data]
        out["new_tokens"] = d.get_new_tokens()
    path = os.path.join(ROOT, "tests", "golden", "dataset_smb.json")
    json.dump(out, open(path, "w"), default=lambda o: o.item() if hasattr(o, "item") else str(o))
    print(f"{path}: {os.path.getsize(path) / 1e3:.1f} kB, {len(out['train'])} train / {len(out['test'])} test samples")


if __name__ == "__main__":
    main()
