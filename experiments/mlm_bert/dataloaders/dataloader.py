"""MLM dataloader (ref. ``experiments/mlm_bert/dataloaders/dataloader.py``): frames are collated with BERT-style
dynamic masking — ``mlm_probability`` of the non-special tokens are selected; 80 % → [MASK], 10 % → random id,
10 % unchanged; unselected labels = −100 (what HF's ``DataCollatorForLanguageModeling`` does, implemented here in
torch so it also works with the offline tokenizer)."""
import torch

from msrflute_b200.core.dataloader import BaseDataLoader
from experiments.mlm_bert.dataloaders.dataset import Dataset


def mlm_collate(features, tokenizer, mlm_probability=0.15, generator=None):
    keys = [k for k in ("input_ids", "attention_mask", "special_tokens_mask", "token_type_ids") if k in features[0]]
    L = max(len(f["input_ids"]) for f in features)
    pad_id = getattr(tokenizer, "pad_token_id", 0) or 0
    fill = {"input_ids": pad_id, "attention_mask": 0, "special_tokens_mask": 1, "token_type_ids": 0}
    batch = {k: torch.tensor([list(f[k]) + [fill[k]] * (L - len(f[k])) for f in features], dtype=torch.long) for k in keys}
    ids = batch["input_ids"]
    special = batch.pop("special_tokens_mask").bool() if "special_tokens_mask" in batch else torch.zeros_like(ids, dtype=torch.bool)
    labels = ids.clone()
    prob = torch.full(ids.shape, mlm_probability).masked_fill(special, 0.0)
    chosen = torch.bernoulli(prob, generator=generator).bool()
    labels[~chosen] = -100
    r = torch.rand(ids.shape, generator=generator)
    mask_id = getattr(tokenizer, "mask_token_id", 3) or 3
    ids = torch.where(chosen & (r < 0.8), torch.full_like(ids, mask_id), ids)
    rnd = torch.randint(len(tokenizer), ids.shape, generator=generator)
    ids = torch.where(chosen & (r >= 0.8) & (r < 0.9), rnd, ids)
    batch["input_ids"], batch["labels"] = ids, labels
    return batch


class DataLoader(BaseDataLoader):
    def __init__(self, mode, data, num_workers=0, **kwargs):
        args = kwargs["args"]
        self.batch_size = args["batch_size"]
        self.mlm_probability = args.get("mlm_probability", 0.15)
        dataset = Dataset(data, args=args, test_only=(mode != "train"), user_idx=kwargs.get("user_idx", None),
                          max_samples_per_user=args.get("max_samples_per_user", -1),
                          min_words_per_utt=args.get("min_words_per_utt", 5))
        self.tokenizer = dataset.tokenizer
        self.vocab_size = len(self.tokenizer)
        super().__init__(dataset, batch_size=self.batch_size, shuffle=(mode == "train"), num_workers=num_workers,
                         collate_fn=lambda feats: mlm_collate(feats, self.tokenizer, self.mlm_probability),
                         drop_last=False)
