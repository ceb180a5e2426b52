"""2-layer LSTM char model for FedShakespeare (ref. ``experiments/nlp_rnn_fedshakespeare/model.py:12-53``):
``Embedding(90, 8, pad=0)`` → ``LSTM(8, 256, layers=2)`` → ``Linear(256, 90)`` at every position; CE with
``ignore_index=0``; accuracy over non-pad targets.  822,570 parameters in 11 tensors."""
import torch
from torch import nn

from ..ops.lstm_ops import LSTM
from ..ops.nn_ops import Embedding
from .common import ClassifierModel


class CharLSTM(nn.Module):
    def __init__(self, embedding_dim=8, vocab_size=90, hidden_size=256, num_layers=2):
        super().__init__()
        self.embeddings = Embedding(vocab_size, embedding_dim, padding_idx=0)      # csrc/nn_kernels.cu gather / scatter-add
        # persistent hand-written LSTM (csrc/lstm_kernels.cu: W_hh resident in shared memory across the 80 steps);
        # same parameter names / shapes / gate order as nn.LSTM, so checkpoints are interchangeable
        self.lstm = LSTM(input_size=embedding_dim, hidden_size=hidden_size, num_layers=num_layers, batch_first=True)
        self.fc = nn.Linear(hidden_size, vocab_size)

    def forward(self, input_seq):
        out, _ = self.lstm(self.embeddings(input_seq.long()))
        return self.fc(out).transpose(1, 2)          # (N, vocab, T)


class RNN(ClassifierModel):
    ignore_index = 0

    def __init__(self, model_config):
        super().__init__(model_config)
        self.net = CharLSTM(embedding_dim=model_config.get("embedding_dim", 8),
                            vocab_size=model_config.get("vocab_size", 90),
                            hidden_size=model_config.get("hidden_size", 256))

    def forward(self, x):
        return self.net(x)
