"""Privacy *attack* metrics evaluated on a client's update (ref. ``extensions/privacy/metrics.py``).

* ``extract_indices_from_embeddings`` (ref :10-22): rows of the embedding gradient with the largest L2 norm
  reveal which token ids occurred in the client's batches; returns (recall, true-positive ids).
* ``practical_epsilon_leakage`` (ref :33-75): per-token perplexity ratio between the received model and the
  model after ONE attacker optimizer step along the client's pseudo-gradient; log of the max (optionally
  probability-weighted) ratio is a practical ε.
"""
import logging
from copy import deepcopy

import numpy as np
import torch as T

from ...utils import make_optimizer, print_rank


def extract_indices_from_embeddings(gradients, batch, embed_size, vocab_size):
    tokens = T.cat([b.reshape(-1) for b in batch]).cpu().numpy()
    embed_grad = gradients[:embed_size * vocab_size].reshape(vocab_size, embed_size)
    valid = tokens[tokens > 0]
    top = T.topk(embed_grad.norm(dim=-1), k=min(len(tokens), vocab_size)).indices.cpu().numpy()
    ratio = float(np.isin(valid, top).mean()) if len(valid) else 0.0
    return ratio, np.intersect1d(top, valid)


def compute_perplexity(encoded_batch, model):
    """Per-token log-likelihood ``[batch, seq]`` of ``encoded_batch`` under ``model``.

    The reference reads ``model.inference(batch)['output']`` as a ``[batch, seq, vocab]`` logit tensor
    (``metrics.py:25-30``) — which none of its shipped models return (the GRU LM returns a dict of top-k
    probabilities or ``None``), so the leakage metric cannot run there.  Here a model may expose
    ``token_logits(batch) -> (logits [b, s, v], targets [b, s], mask [b, s])``; otherwise a 3-D ``output`` tensor is
    used the reference's way."""
    if hasattr(model, "token_logits"):
        logits, targets, mask = model.token_logits(encoded_batch)
        logp = T.nn.functional.log_softmax(logits.float(), dim=-1)
        ll = logp.gather(-1, targets.clamp(min=0).long().unsqueeze(-1)).squeeze(-1)
        return T.where(mask, ll, T.zeros_like(ll))
    out = model.inference(encoded_batch)["output"]
    if not T.is_tensor(out) or out.dim() != 3:
        raise TypeError("privacy leakage metric needs per-token logits: give the model a token_logits(batch) method")
    b, s, v = out.shape
    logp = T.nn.functional.log_softmax(out, dim=-1)
    tgt = encoded_batch.reshape(b, s, 1).to(logp.device).long()
    return logp.gather(-1, tgt).squeeze(-1)


def practical_epsilon_leakage(original_params, model, encoded_batches, is_weighted_leakage=True, max_ratio=1e9,
                              optimizer_config=None):
    current_params = deepcopy(model.state_dict())
    current_grads = {n: p.grad.clone().detach() for n, p in model.named_parameters()}
    with T.no_grad():
        for (n, p) in model.named_parameters():
            p.data.copy_(original_params[n].to(p.device))
    tol = 1 / max_ratio
    max_leak = 0.0
    with T.no_grad():
        pre = [compute_perplexity(b, model) for b in encoded_batches]
        for n, p in model.named_parameters():
            p.grad = current_grads[n].clone()
        make_optimizer(optimizer_config or {"lr": 0.03, "amsgrad": False, "type": "adamax"}, model).step()
        post = [compute_perplexity(b, model) for b in encoded_batches]
        for a, b in zip(pre, post):
            leak = ((a + tol) / (b + tol)).clamp_(0, max_ratio)
            if is_weighted_leakage:
                leak = T.max(a.exp(), b.exp()) * leak
            max_leak = max(max_leak, leak.max().item())
    print_rank("raw max leakage: {}".format(max_leak), loglevel=logging.DEBUG)
    model.load_state_dict(current_params)
    from ...parallel.arena import rebind_grads
    ar = getattr(model, "_flute_arena", None)
    if ar is not None and ar[1] is not None:
        rebind_grads(model)
        for (n, p) in model.named_parameters():
            p.grad.copy_(current_grads[n])
    else:
        for n, p in model.named_parameters():
            p.grad = current_grads[n]
    return max(float(np.log(max_leak)) if max_leak > 0 else 0.0, 0)
