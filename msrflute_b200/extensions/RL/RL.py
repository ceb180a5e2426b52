"""Learned aggregation weights: a small Q-network maps the round's client statistics
``[weights ‖ grad magnitudes ‖ grad means ‖ grad variances]`` to per-client log-weights
(ref. ``extensions/RL/RL.py``: networks :14-144, agent :149-343).

Behavioural parity: ε-greedy exploration with multiplicative annealing, bounded replay memory, MSE between
``Σ(Q·action)`` and the reward ∈ {−1, 0.1, 1}, optional BiLSTM trunk over the last ``minibatch_size`` states,
checkpoint files ``rl_<K>.<descriptor>.model`` / ``.stats`` under ``RL.RL_path``.

The reference's ``make_model`` moves an undefined ``model`` to the device (``RL.py:274``) and so cannot be
constructed; this implementation is the working equivalent of the evident intent.
"""
import json
import logging
import os
import random
from collections import OrderedDict

import numpy as np
import torch
import torch.nn as nn

from ...utils import make_lr_scheduler, make_optimizer, print_rank, to_device, torch_save, try_except_save


class SequenceWise(nn.Module):
    """Apply ``module`` to a (T, N, H) tensor by folding T and N together."""

    def __init__(self, module):
        super().__init__()
        self.module = module

    def forward(self, x):
        t, n = x.shape[0], x.shape[1]
        return self.module(x.reshape(t * n, -1)).view(t, n, -1)


class BatchRNN(nn.Module):
    def __init__(self, input_size, hidden_size, rnn_type=nn.LSTM, bidirectional=False, batch_norm=True,
                 dropout=0.0, multi=1):
        super().__init__()
        self.bidirectional, self.multi = bidirectional, multi
        self.batch_norm = SequenceWise(nn.BatchNorm1d(input_size)) if batch_norm else None
        self.rnn = rnn_type(input_size=input_size, hidden_size=hidden_size, bidirectional=bidirectional, bias=True,
                            batch_first=True, dropout=dropout)

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(1)
        if self.batch_norm is not None:
            x = self.batch_norm(x.contiguous())
        x, _ = self.rnn(x)
        if self.bidirectional and self.multi < 2:      # sum the two directions
            x = x.view(x.size(0), x.size(1), 2, -1).sum(2)
        return x


class NeuralNetwork(nn.Module):
    """MLP (``"in,h1,…,out"``) or BiLSTM stack + 2-layer head."""

    def __init__(self, params, wantLSTM=False, batch_norm=False):
        super().__init__()
        dims = [int(x) for x in str(params).split(",")]
        self.wantLSTM = wantLSTM
        self.input_norm = None
        if wantLSTM:
            rnns = [(str(i - 1), BatchRNN(dims[i - 1], dims[i], nn.LSTM, bidirectional=True, batch_norm=batch_norm))
                    for i in range(1, len(dims) - 2)]
            self.rnn = nn.Sequential(OrderedDict(rnns))
            head = nn.Sequential(nn.Linear(dims[-3], dims[-2]), nn.ReLU(inplace=True), nn.Linear(dims[-2], dims[-1]))
            self.mlp = SequenceWise(head)
        else:
            if batch_norm:
                self.input_norm = nn.BatchNorm1d(dims[0])
            layers = []
            for i in range(1, len(dims) - 1):
                layers += [nn.Linear(dims[i - 1], dims[i]), nn.ReLU(inplace=True)]
            layers.append(nn.Linear(dims[-2], dims[-1]))
            self.mlp = nn.Sequential(*layers)

    def forward(self, x):
        if self.wantLSTM:
            x = self.rnn(x)
        elif self.input_norm is not None:
            x = self.input_norm(x if x.dim() > 1 else x.unsqueeze(0))
        return self.mlp(x).squeeze()


class RL:
    def __init__(self, config=None, model_path=None):
        self.config = config
        rl = config["RL"]
        self.out_size = config["num_clients_per_iteration"]
        self.wantLSTM = rl.get("wantLSTM", False)
        self.replay_memory, self.state_memory = [], []
        self.epsilon = rl["initial_epsilon"]
        self.step, self.runningLoss = 0, 0
        desc = rl.get("model_descriptor_RL", "Default")
        # the reference requires RL.RL_path (KeyError otherwise, RL.py:164); default to RL_models next to the checkpoints
        # of THIS run (never the working directory: runs must not write into the source tree)
        base = model_path or config.get("model_path", None)
        if base is None:
            import tempfile
            base = tempfile.mkdtemp(prefix="flute_rl_")
        rl_path = rl.get("RL_path") or os.path.join(base, "RL_models")
        os.makedirs(rl_path, exist_ok=True)
        self.model_name = os.path.join(rl_path, "rl_{}.{}.model".format(self.out_size, desc))
        self.stats_name = os.path.join(rl_path, "rl_{}.{}.stats".format(self.out_size, desc))
        self.make_model()
        self.load_saved_status()
        self.rl_weights = None
        self.rl_losses = [None, None]
        self.criterion = nn.MSELoss()

    def set_losses(self, losses):
        self.rl_losses = losses

    def set_weights(self, weights):
        self.rl_weights = weights

    def forward(self, state=None):
        state = np.asarray(state, dtype=np.float64)
        if self.wantLSTM:
            state = state.reshape(1, -1)
            if len(self.state_memory) == 0:
                self.state_memory = np.zeros((self.config["RL"]["minibatch_size"], state.shape[1]))
            self.state_memory = np.concatenate((self.state_memory[1:], state), axis=0)
            state = self.state_memory
        if random.random() <= self.epsilon:
            print_rank("Performed random action!", logging.DEBUG)
            return to_device(torch.rand(self.out_size))
        self.model.eval()
        with torch.no_grad():
            return self.model(to_device(torch.from_numpy(np.ascontiguousarray(state))).float())

    def train(self, batch=None):
        rl = self.config["RL"]
        self.replay_memory.append(batch)
        if len(self.replay_memory) > rl["max_replay_memory_size"]:
            self.replay_memory.pop(0)
        if self.epsilon * rl["epsilon_gamma"] > rl["final_epsilon"]:
            self.epsilon *= rl["epsilon_gamma"]
        mb = rl["minibatch_size"]
        minibatch = self.replay_memory[-mb:] if self.wantLSTM else \
            random.sample(self.replay_memory, min(len(self.replay_memory), mb))
        state = to_device(torch.tensor(np.stack([np.asarray(d[0], dtype=np.float32) for d in minibatch])))
        action = to_device(torch.tensor(np.stack([np.asarray(d[1], dtype=np.float32) for d in minibatch])))
        reward = to_device(torch.tensor(np.asarray([d[2] for d in minibatch], dtype=np.float32))).reshape(-1)
        self.model.train()
        if state.shape[0] == 1 and any(isinstance(m, nn.BatchNorm1d) for m in self.model.modules()):
            self.model.eval()           # BatchNorm cannot train on a single sample
        out = self.model(state)
        out = out.reshape(state.shape[0], -1)[:, :action.shape[1]]
        q = torch.sum(out * action, dim=1)
        self.optimizer.zero_grad()
        loss = self.criterion(q, reward.detach())
        loss.backward()
        self.optimizer.step()
        lv = loss.item()
        self.runningLoss = lv if self.runningLoss == 0 else 0.95 * self.runningLoss + 0.05 * lv
        print_rank("Running Loss for RL training process: {}".format(self.runningLoss), logging.DEBUG)
        self.lr_scheduler.step()

    def make_model(self):
        rl = self.config["RL"]
        self.model = to_device(NeuralNetwork(rl["network_params"], rl.get("wantLSTM", False), rl.get("batchNorm", False)))
        self.optimizer = make_optimizer(rl["optimizer_config"], self.model)
        self.lr_scheduler = make_lr_scheduler(rl["annealing_config"], self.optimizer, num_batches=1)

    def load_saved_status(self):
        if os.path.exists(self.model_name):
            print_rank("Resuming from checkpoint model {}".format(self.model_name))
            self.load()
        if os.path.exists(self.stats_name):
            with open(self.stats_name) as f:
                e = json.load(f)
            self.cur_iter_no, self.val_loss, self.val_cer = e["i"], e["val_loss"], e["val_cer"]
            self.runningLoss = e["weight"]

    def load(self):
        ckpt = torch.load(self.model_name, map_location=next(self.model.parameters()).device, weights_only=False)
        self.model.load_state_dict(ckpt["model_state_dict"])
        if self.optimizer is not None and ckpt.get("optimizer_state_dict"):
            self.optimizer.load_state_dict(ckpt["optimizer_state_dict"])
        sd = ckpt.get("lr_scheduler_state_dict")
        if sd and self.lr_scheduler is not None:
            self.lr_scheduler.load_state_dict(sd)

    def save(self, i):
        state = {"model_state_dict": self.model.state_dict(),
                 "optimizer_state_dict": self.optimizer.state_dict() if self.optimizer is not None else None,
                 "lr_scheduler_state_dict": self.lr_scheduler.state_dict() if self.lr_scheduler is not None else None}
        os.makedirs(os.path.dirname(self.model_name) or ".", exist_ok=True)
        try_except_save(torch_save, state_or_model=state, save_path=self.model_name)
        losses = self.rl_losses if self.rl_losses and None not in self.rl_losses else (float("nan"), float("nan"))
        with open(self.stats_name, "w") as f:
            json.dump({"i": i + 1, "val_loss": float(losses[0]), "val_cer": float(losses[1]),
                       "weight": float(self.runningLoss)}, f)
