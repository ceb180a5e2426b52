"""FedNewsRec (Qi et al., EMNLP-Findings 2020) — ref. ``experiments/fednewsrec/{fednewsrec_model,model,utils}.py``.

Architecture parity: frozen (GloVe) title-word embedding → DocEncoder [dropout → Conv1d(300→400, k=3) → ReLU →
dropout → 20-head × 20-dim self-attention → ReLU → dropout → attentive pooling] ; UserEncoder [self-attention over
the clicked-news vectors → attentive pooling] ⊕ [GRU(400) over the last 20 clicks] → attentive pooling of the two;
score = ⟨candidate vec, user vec⟩; training loss = CE over (1 positive + ``npratio`` negatives); evaluation metrics
AUC / MRR / nDCG@5 / nDCG@10 per impression.

B200-first: the reference applies the doc encoder to each of the 50 clicked + 5 candidate titles in a python loop
growing a tensor with ``torch.cat`` (``TimeDistributed``, ref :193-206) and uses einsum attention; here all titles of a
batch go through the encoder as ONE batch and attention is ``scaled_dot_product_attention`` (flash kernels).  The
ranking metrics are computed on the device (no sklearn / numpy round trip per impression).
"""
import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from ..core.model import BaseModel

npratio = 4


class AttentivePooling(nn.Module):
    def __init__(self, dim1: int, dim2: int):
        super().__init__()
        self.dropout = nn.Dropout(0.2)
        self.dense = nn.Linear(dim2, 200)
        self.dense2 = nn.Linear(200, 1)

    def forward(self, x):
        v = self.dropout(x)
        att = torch.softmax(self.dense2(torch.tanh(self.dense(v))).squeeze(-1), dim=1)
        return torch.einsum("ijk,ij->ik", v, att)


class Attention(nn.Module):
    def __init__(self, input_dim, nb_head, size_per_head):
        super().__init__()
        self.nb_head, self.size_per_head, self.output_dim = nb_head, size_per_head, nb_head * size_per_head
        self.WQ = nn.Linear(input_dim, self.output_dim, bias=False)
        self.WK = nn.Linear(input_dim, self.output_dim, bias=False)
        self.WV = nn.Linear(input_dim, self.output_dim, bias=False)
        for l in (self.WQ, self.WK, self.WV):
            nn.init.xavier_uniform_(l.weight, gain=np.sqrt(2))

    def forward(self, x):
        q_in, k_in, v_in = x if isinstance(x, (list, tuple)) else (x, x, x)
        split = lambda t: t.view(t.shape[0], t.shape[1], self.nb_head, self.size_per_head).transpose(1, 2)
        o = F.scaled_dot_product_attention(split(self.WQ(q_in)), split(self.WK(k_in)), split(self.WV(v_in)))
        return o.transpose(1, 2).reshape(q_in.shape[0], q_in.shape[1], self.output_dim)


class DocEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.drop1, self.conv, self.drop2 = nn.Dropout(0.2), nn.Conv1d(300, 400, 3), nn.Dropout(0.2)
        self.attention = Attention(400, 20, 20)
        self.drop3, self.pool = nn.Dropout(0.2), AttentivePooling(30, 400)

    def forward(self, x):                                    # x: [N, words, 300]
        h = self.drop2(F.relu(self.conv(self.drop1(x).transpose(-2, -1)))).transpose(-2, -1)
        return self.pool(self.drop3(F.relu(self.attention(h))))


class UserEncoder(nn.Module):
    def __init__(self):
        super().__init__()
        self.attention2 = Attention(400, 20, 20)
        self.dropout2 = nn.Dropout(0.2)
        self.pool2 = AttentivePooling(50, 400)
        self.gru2 = nn.GRU(400, 400, bidirectional=False, batch_first=True)
        self.pool3 = AttentivePooling(2, 400)

    def forward(self, news_vecs):                            # [B, clicks, 400]
        long_term = self.pool2(self.dropout2(self.attention2(news_vecs)))
        short_term = self.gru2(news_vecs[:, -20:, :])[0][:, -1, :]
        return self.pool3(torch.stack([short_term, long_term], dim=1))


class FedNewsRec(nn.Module):
    def __init__(self, title_word_embedding_matrix):
        super().__init__()
        self.doc_encoder = DocEncoder()
        self.user_encoder = UserEncoder()
        self.title_word_embedding_layer = nn.Embedding.from_pretrained(
            torch.as_tensor(np.asarray(title_word_embedding_matrix), dtype=torch.float), freeze=True)

    def news_encoder(self, news_title):                      # [N, words] → [N, 400]
        return self.doc_encoder(self.title_word_embedding_layer(news_title))

    def _encode_many(self, titles):                          # [B, n, words] → [B, n, 400]  (one batched pass)
        B, n, w = titles.shape
        return self.news_encoder(titles.reshape(B * n, w)).view(B, n, -1)

    def forward(self, click_title, can_title):
        user_vec = self.user_encoder(self._encode_many(click_title))
        scores = torch.einsum("ijk,ik->ij", self._encode_many(can_title), user_vec)
        return scores, user_vec


# ---------------------------------------------------------------------------- ranking metrics (device tensors)
def dcg_score(y_true, y_score, k=10):
    order = torch.argsort(y_score, descending=True)[:k]
    gains = 2.0 ** y_true[order].float() - 1
    return (gains / torch.log2(torch.arange(len(order), device=y_true.device).float() + 2)).sum()


def ndcg_score(y_true, y_score, k=10):
    return dcg_score(y_true, y_score, k) / dcg_score(y_true, y_true.float(), k).clamp(min=1e-12)


def mrr_score(y_true, y_score):
    order = torch.argsort(y_score, descending=True)
    yt = y_true[order].float()
    return (yt / (torch.arange(len(yt), device=yt.device).float() + 1)).sum() / yt.sum().clamp(min=1e-12)


def auc_score(y_true, y_score):
    """Mann-Whitney AUC with tie handling (== sklearn.metrics.roc_auc_score for binary labels)."""
    pos, neg = y_score[y_true > 0], y_score[y_true <= 0]
    if pos.numel() == 0 or neg.numel() == 0:
        return torch.tensor(0.5, device=y_score.device)
    diff = pos.view(-1, 1) - neg.view(1, -1)
    return ((diff > 0).float() + 0.5 * (diff == 0).float()).mean()


class FEDNEWS(BaseModel):
    """``model_config``: ``embbeding_path`` (sic — the reference's key) with MIND ``train|val/news.tsv`` +
    ``glove.840B.300d.txt``; without it (or with ``synthetic_vocab``) a random frozen embedding table is used."""

    def __init__(self, model_config):
        super().__init__()
        path = model_config.get("embbeding_path", None)
        matrix = None
        if path and path != "None":
            try:
                from experiments.fednewsrec.dataloaders.preprocess_mind import load_matrix, read_news
                _, _, _, _, word_dict = read_news(path, ["train", "val"])
                matrix, _ = load_matrix(path, word_dict)
            except (FileNotFoundError, ImportError):
                matrix = None
        if matrix is None:
            vocab = int(model_config.get("synthetic_vocab", 5000))
            matrix = np.random.default_rng(0).standard_normal((vocab + 1, 300)).astype(np.float32) * 0.3
        self.net = FedNewsRec(matrix)

    def _dev(self):
        return self.net.doc_encoder.conv.weight.device

    def loss(self, input):
        if not self.net.training:
            return torch.zeros((), device=self._dev())            # the loss is not used during evaluation
        (click, sample), label = input["x"], input["y"]
        dev = self._dev()
        out, _ = self.net(click.to(dev), sample.to(dev))
        return F.cross_entropy(out, label.to(dev).long())

    def inference(self, input):
        (hist, imp), labels = input["x"], input["y"]
        dev = self._dev()
        with torch.no_grad():
            hist, imp = hist.to(dev), imp.to(dev)
            if hist.dim() == 3:
                hist, imp, labels = hist[0], imp[0], labels[0]
            nv = self.net.news_encoder(imp)
            uv = self.net.user_encoder(self.net.news_encoder(hist).unsqueeze(0))[0]
            score = nv @ uv
            y = torch.as_tensor(labels, device=dev).float().reshape(-1)
            m = {"auc": auc_score(y, score), "mrr": mrr_score(y, score), "ndcg1": ndcg_score(y, score, 1),
                 "ndcg5": ndcg_score(y, score, 5), "ndcg10": ndcg_score(y, score, 10)}
        return {"output": None, "acc": m["ndcg1"].item(), "batch_size": 1,
                "auc": {"value": m["auc"].item(), "higher_is_better": True},
                "mrr": {"value": m["mrr"].item(), "higher_is_better": True},
                "ndcg5": {"value": m["ndcg5"].item(), "higher_is_better": True},
                "ndcg10": {"value": m["ndcg10"].item(), "higher_is_better": True}}
