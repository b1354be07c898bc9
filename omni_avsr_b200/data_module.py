"""Length-bucketed frame-budget batching and the batch collation of the reference's input pipeline
(/root/reference/datamodule/data_module.py), host side.

  * `batch_by_token_count`  (:82-98)  greedy packing of (index, length) pairs under a frame budget (+ optional size cap);
  * `CustomBucketDataset`   (:101-140) lengths -> `num_buckets` equal-width buckets (torch.linspace + torch.bucketize),
    items ordered longest-first (or shuffled with Python's `random.sample`, the reference's RNG) and then stably by
    bucket, packed with the frame budget: every batch holds clips of similar length whose total is <= max_frames
    (README recipes: --max-frames-audiovisual 1500 video frames, i.e. 60 s of 25 fps video per batch and GPU);
  * `collate_LLM`           (:19-79)  media zero-padded to the batch maximum, text tokenised with padding='longest',
    labels = tokens with <pad> -> -100, `lengths` = audio sample counts; inference: tokens = [[bos]] / [[]] + gold_text.

These feed the drop-in ModelModule_LLM with the variable-length batches the reference trains on (instead of the fixed
16 s synthetic clips of the headline benchmark); the kernels behind it take any (B, T) -- the token count follows
max(int(max_len / 16000 * 50), 25) and the padded text / causal-mask semantics of modeling_OmniAVSR.py:537.
File decoding, the dataset class and the DataLoader workers stay out of scope (dataloader)."""
from __future__ import annotations

import random
from typing import List, Optional, Sequence, Tuple

import torch

IGNORE_INDEX = -100


def batch_by_token_count(idx_target_lengths: Sequence[Tuple[int, int]], max_frames: int,
                         batch_size: Optional[int] = None) -> List[List[int]]:
    batches: List[List[int]] = []
    current: List[int] = []
    count = 0
    for idx, length in idx_target_lengths:
        if count + length > max_frames or (batch_size and len(current) == batch_size):
            batches.append(current)            # (an empty first batch when the very first clip exceeds the budget: as :88-91)
            current, count = [idx], length
        else:
            current.append(idx)
            count += length
    if current:
        batches.append(current)
    return batches


class CustomBucketDataset(torch.utils.data.Dataset):
    """dataset[i] -> the list of samples of batch i (the DataLoader runs with batch_size=None)."""

    def __init__(self, dataset, lengths, max_frames, num_buckets, shuffle=False, batch_size=None):
        super().__init__()
        assert len(dataset) == len(lengths)
        self.dataset = dataset
        max_length, min_length = max(lengths), min(lengths)
        assert max_frames >= max_length
        buckets = torch.linspace(min_length, max_length, num_buckets)
        lengths_t = torch.tensor(lengths)
        assignments = torch.bucketize(lengths_t, buckets)
        items = [(idx, length, assignments[idx]) for idx, length in enumerate(lengths_t)]
        if shuffle:
            # what :125-128 evidently intends; in the reference itself this branch raises NameError (`random` is never
            # imported in data_module.py) and is never taken: train_dataloader shuffles the batches, not the items
            items = random.sample(items, len(items))
        else:
            items = sorted(items, key=lambda x: x[1], reverse=True)
        items = sorted(items, key=lambda x: x[2])                   # stable: keeps the order above inside a bucket
        self.batches = batch_by_token_count([(int(i), int(l)) for i, l, _ in items], max_frames, batch_size=batch_size)

    def __getitem__(self, idx):
        return [self.dataset[sub] for sub in self.batches[idx]]

    def __len__(self):
        return len(self.batches)


def collate_LLM(batch, tokenizer, modality, is_trainval=True):
    """Same batch dict as the reference's collate (keys tokens / labels / audio / lengths / video [/ gold_text])."""
    pad_id = tokenizer.convert_tokens_to_ids("<pad>") if not getattr(tokenizer, "is_qwen", False) else tokenizer.pad_token_id
    has_a = modality in ("audio", "audiovisual", "audiovisual_avhubert")
    has_v = modality in ("video", "audiovisual", "audiovisual_avhubert")
    out = {}
    lengths = []
    if is_trainval:
        texts = [b["tokens"] for b in batch]
        audios = [b["audio"] for b in batch] if has_a else None
        videos = [b["video"] for b in batch] if has_v else None
        if has_a:
            lengths = [len(a) for a in audios]
        tokens = tokenizer(texts, padding="longest", return_tensors="pt").input_ids
        labels = torch.where(tokens == pad_id, torch.full_like(tokens, IGNORE_INDEX), tokens)
        out["tokens"], out["labels"] = tokens, labels
        if has_a:
            out["audio"] = torch.nn.utils.rnn.pad_sequence(audios, batch_first=True, padding_value=0)
            out["lengths"] = torch.tensor(lengths)
        if has_v:
            out["video"] = torch.nn.utils.rnn.pad_sequence(videos, batch_first=True, padding_value=0)
        return out
    # inference: one utterance, only the BOS token (Llama) or no token at all (Qwen); the transcript travels as gold_text
    if getattr(tokenizer, "is_qwen", False) or "Qwen" in str(getattr(tokenizer, "name_or_path", "")):
        tokens = torch.tensor([[]], dtype=torch.long)
    else:
        tokens = torch.tensor([tokenizer.vocab["<|begin_of_text|>"]]).unsqueeze(0)
    out["gold_text"] = batch["tokens"]
    out["tokens"], out["labels"] = tokens, None
    if has_a:
        out["audio"] = batch["audio"].unsqueeze(0)
        out["lengths"] = torch.tensor([len(batch["audio"])])
    if has_v:
        out["video"] = batch["video"].unsqueeze(0)
    return out


class SyntheticLengthDataset(torch.utils.data.Dataset):
    """Synthetic LRS3-shaped utterances with a given list of video-frame counts (25 fps; audio = 640 samples per frame):
    what AVDataset_LLM yields after the transforms, without files.  Used by the bucketing tests and `bench.py --workload
    ragged`."""

    def __init__(self, frame_counts: Sequence[int], text_words=(4, 24), seed=0):
        self.frames = [int(f) for f in frame_counts]
        self.input_lengths = list(self.frames)
        self.seed = seed
        self.text_words = text_words

    def __len__(self):
        return len(self.frames)

    def __getitem__(self, i):
        g = torch.Generator().manual_seed(self.seed * 1000003 + i)
        T = self.frames[i]
        n_words = int(torch.randint(self.text_words[0], self.text_words[1] + 1, (1,), generator=g))
        words = " ".join(f"w{int(w)}" for w in torch.randint(0, 5000, (n_words,), generator=g))
        audio = torch.randn(T * 640, 1, generator=g)
        audio = torch.nn.functional.layer_norm(audio, audio.shape)
        video = (torch.rand(T, 1, 88, 88, generator=g) - 0.421) / 0.165
        return {"tokens": words, "audio": audio, "video": video}
