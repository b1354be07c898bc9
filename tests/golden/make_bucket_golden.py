"""Generates tests/golden/bucket_golden.json by EXECUTING the reference's own datamodule/data_module.py
(`_batch_by_token_count`, `CustomBucketDataset`, `collate_LLM`) from /root/reference on seeded inputs.
The module imports pytorch_lightning and its sibling dataset / transform modules at import time; those are stubbed by name
(nothing of them is used by the three functions).  Run in the build container:  python tests/golden/make_bucket_golden.py"""
import importlib.util
import json
import os
import random
import sys
import types

import torch

REF = "/root/reference/datamodule/data_module.py"
HERE = os.path.dirname(os.path.abspath(__file__))


def load_reference():
    pl = types.ModuleType("pytorch_lightning")
    pl.LightningDataModule = object
    sys.modules.setdefault("pytorch_lightning", pl)
    pkg = types.ModuleType("datamodule")
    pkg.__path__ = []
    sys.modules["datamodule"] = pkg
    for name, attrs in (("av_dataset", ["AVDataset_LLM"]), ("transforms", ["AudioTransform", "VideoTransform"])):
        m = types.ModuleType("datamodule." + name)
        for a in attrs:
            setattr(m, a, object)
        sys.modules["datamodule." + name] = m
    spec = importlib.util.spec_from_file_location("datamodule.data_module", REF)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["datamodule.data_module"] = mod
    spec.loader.exec_module(mod)
    return mod


class Tok:
    """Minimal tokenizer with the attributes collate_LLM touches (Llama-3.2-1B branch)."""
    name_or_path = "meta-llama/Llama-3.2-1B"
    vocab = {"<|begin_of_text|>": 128000}

    def convert_tokens_to_ids(self, t):
        return 128256

    def __call__(self, texts, padding=None, return_tensors=None):
        rows = [[128000] + [1000 + len(w) for w in t.split()] + [128001] for t in texts]
        L = max(len(r) for r in rows)

        class R:
            pass
        r = R()
        r.input_ids = torch.tensor([x + [128256] * (L - len(x)) for x in rows])
        return r


def main():
    ref = load_reference()
    out = {"cases": []}
    g = torch.Generator().manual_seed(0)
    for n, lo, hi, max_frames, buckets, bs, shuffle, seed in (
            (200, 20, 400, 1500, 50, None, False, 0), (500, 20, 400, 1500, 50, None, False, 1),
            (57, 50, 155, 600, 8, 4, False, 2), (300, 25, 600, 1000, 1, None, False, 3), (40, 30, 31, 100, 5, 3, False, 4)):
        # (shuffle=True is dead code in the reference: data_module.py never imports `random`, so :126 raises NameError; its
        # train_dataloader leaves it False and shuffles the BATCHES with DataLoader(shuffle=True) instead)
        lengths = torch.randint(lo, hi + 1, (n,), generator=g).tolist()
        random.seed(seed)
        ds = ref.CustomBucketDataset(list(range(n)), lengths, max_frames, buckets, shuffle=shuffle, batch_size=bs)
        out["cases"].append(dict(lengths=lengths, max_frames=max_frames, num_buckets=buckets, batch_size=bs, shuffle=shuffle,
                                 seed=seed, batches=[[int(i) for i in b] for b in ds.batches]))
    pairs = [(i, l) for i, l in enumerate([5, 9, 3, 12, 1, 7, 7, 2])]
    out["token_count"] = [dict(pairs=pairs, max_frames=mf, batch_size=bs,
                               batches=ref._batch_by_token_count(pairs, mf, batch_size=bs))
                          for mf, bs in ((12, None), (20, 2), (4, None))]
    # collate: three ragged utterances
    items = []
    for i, (T, text) in enumerate(((3, "a bb ccc"), (5, "dddd"), (2, "e ff g hh iiii"))):
        items.append({"tokens": text, "audio": torch.arange(T * 4, dtype=torch.float32).view(-1, 1) + i,
                      "video": torch.arange(T * 2, dtype=torch.float32).view(T, 1, 1, 2) + i})
    tr = ref.collate_LLM(items, Tok(), "audiovisual", is_trainval=True)
    te = ref.collate_LLM(items[1], Tok(), "audiovisual", is_trainval=False)
    out["collate_train"] = {k: v.tolist() for k, v in tr.items()}
    out["collate_test"] = {k: (v.tolist() if torch.is_tensor(v) else v) for k, v in te.items()}
    json.dump(out, open(os.path.join(HERE, "bucket_golden.json"), "w"))
    print("wrote bucket_golden.json:", [len(c["batches"]) for c in out["cases"]], "batches per case")


if __name__ == "__main__":
    main()
