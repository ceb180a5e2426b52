"""General helpers: logging, IO with retry, optimizer/scheduler factories, misc math.

Parity target: ``utils/utils.py`` of the reference (``make_optimizer`` :27,
``make_lr_scheduler`` :151, schedulers :189-294, ``init_logging`` :299,
``print_rank`` :319, ``try_except_save`` :348, ``update_json_log`` :546,
``scrub_empty_clients`` :563, ``compute_grad_cosines`` :585,
``convex_inference`` :598, ``alpha_update`` :605, ``get_label_VAT`` :620).

B200-first differences: ``print_rank`` really is rank-tagged; tensors are moved
with ``to_device`` onto the *current* CUDA device (one process per GPU); the
gradient helpers operate on the flat arena when one is attached (single fused
reduction, no per-tensor ``.item()``).
"""
from __future__ import annotations

import copy
import io
import json
import logging
import math
import os
import pstats
import sys
import time
from collections import OrderedDict

import numpy as np
import torch
import yaml

from .optimizers import AdamW, LAMB, LarsSGD, LarsSGDV1, LARSWrapper


# --------------------------------------------------------------------- env
def env_rank() -> int:
    return int(os.environ.get("RANK", 0))


def env_local_rank() -> int:
    return int(os.environ.get("LOCAL_RANK", 0))


def env_world_size() -> int:
    return int(os.environ.get("WORLD_SIZE", 1))


def default_device() -> torch.device:
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def to_device(x):
    """Move a tensor/module to this rank's device (ref. ``utils.py:543``)."""
    return x.to(default_device()) if torch.cuda.is_available() else x


# ----------------------------------------------------------------- logging
def init_logging(log_dir, loglevel=logging.DEBUG):
    os.makedirs(log_dir, exist_ok=True)
    root = logging.getLogger()
    for h in list(root.handlers):
        root.removeHandler(h)
    root.setLevel(loglevel)
    fh = logging.FileHandler(os.path.join(log_dir, "log.out"))
    root.addHandler(fh)
    root.addHandler(logging.StreamHandler(stream=sys.stdout))


def print_rank(msg, loglevel=logging.INFO):
    logging.log(loglevel, "{} [r{}] : {}".format(time.ctime(), env_rank(), msg))


def print_profiler(profiler, loglevel=logging.INFO, top=20):
    buf = io.StringIO()
    pstats.Stats(profiler, stream=buf).strip_dirs().sort_stats(pstats.SortKey.CUMULATIVE).print_stats(top)
    for line in buf.getvalue().split("\n"):
        print_rank(line, loglevel=loglevel)


def print_cuda_stats():
    if torch.cuda.is_available():
        print_rank("cuda mem allocated={} reserved={}".format(
            torch.cuda.memory_allocated(), torch.cuda.memory_reserved()))
    else:
        print_rank("No CUDA GPU available")


# ---------------------------------------------------------------------- IO
_YAML_WRITTEN = {}


def write_yaml(save_path, config):
    """The reference rewrites ``config.yaml`` next to every checkpoint (every round); dumping it costs ~5 ms of host
    time, so an unchanged config is not written again."""
    if hasattr(config, "to_dict"):
        config = config.to_dict()
    key = repr(config)
    if _YAML_WRITTEN.get(save_path) == key and os.path.exists(save_path):
        return
    with open(save_path, "w", encoding="utf8") as f:
        yaml.safe_dump(config, f, default_flow_style=False)
    _YAML_WRITTEN[save_path] = key


def torch_save(save_path, state_or_model):
    tmp = save_path + ".tmp"
    torch.save(state_or_model, tmp)
    os.replace(tmp, save_path)     # atomic: a crash never leaves a torn checkpoint


def write_tokens(save_path, token_list):
    with open(save_path, "w", encoding="utf8") as f:
        for w in token_list:
            f.write(w + "\n")


def try_except_save(save_fn, max_attempts=3, **kwargs):
    """Retry a write up to three times on IOError (ref. ``utils.py:348-359``)."""
    for attempt in range(1, max_attempts + 1):
        try:
            save_fn(**kwargs)
        except IOError:
            print_rank("Write operation failed on {} attempt".format(attempt))
        else:
            print_rank("Write operation succeeded in {} attempts".format(attempt), logging.DEBUG)
            return True
    return False


_JSON_LOGS = {}


def update_json_log(log_path, status_info, background=False):
    """Merge ``status_info`` into the JSON file at ``log_path`` (ref. ``utils.py:361-377``).  ``background=True``
    hands the write to the async checkpoint thread (latest wins, durable after ``flush_checkpoints()``) — two file
    operations per round otherwise sit on the round's critical path."""
    elems = _JSON_LOGS.get(log_path) if background else None
    if elems is None:
        elems = {}
        if os.path.exists(log_path):
            with open(log_path, "r") as f:
                elems = json.load(f)
    elems.update(status_info)
    if background:
        from .async_ckpt import get_checkpointer
        _JSON_LOGS[log_path] = elems
        get_checkpointer().submit_text(log_path, json.dumps(elems))
        return
    if _JSON_LOGS.pop(log_path, None) is not None:      # a background write of this file may still be queued
        from .async_ckpt import flush_checkpoints
        flush_checkpoints()
        with open(log_path, "r") as f:
            elems = json.load(f)
        elems.update(status_info)
    tmp = log_path + ".tmp"
    with open(tmp, "w") as f:
        json.dump(elems, f)
    os.replace(tmp, log_path)


def find_pretrained_model(model_path, config):
    out = config.get("pretrained_model_path", None) if config is not None else None
    print_rank("Loading Model from: {}".format(out), loglevel=logging.INFO)
    return out


# -------------------------------------------------------------- optimizers
def make_optimizer(optimizer_config, model_or_params):
    """Build one of sgd/adam/adamax/lars/LarsSGD/lamb/adamW (ref. ``utils.py:27-64``)."""
    cfg = dict(copy.deepcopy(optimizer_config))
    kind = cfg.pop("type")
    params = model_or_params.parameters() if hasattr(model_or_params, "parameters") else model_or_params
    if kind == "sgd":
        return torch.optim.SGD(params, **cfg)
    if kind == "adam":
        return torch.optim.Adam(params, **cfg)
    if kind == "adamax":
        cfg.pop("amsgrad", None)
        return torch.optim.Adamax(params, **cfg)
    if kind == "lars":
        # the reference wraps SGD with the external ``torchlars`` package
        # (eps=1e-8, trust_coef=0.001); LARSWrapper is a native equivalent.
        return LARSWrapper(torch.optim.SGD(params, **cfg), eps=1e-8, trust_coef=0.001)
    if kind == "LarsSGD":
        return LarsSGD(params, **cfg)
    if kind == "lamb":
        return LAMB(params, **cfg)
    if kind == "adamW":
        cfg.pop("amsgrad", None)
        return AdamW(params, **cfg)
    raise ValueError("{} optimizer not supported".format(kind))


def get_lr(optimizer):
    for g in optimizer.param_groups:
        return g["lr"]


def get_lr_all(optimizer):
    for g in optimizer.param_groups:
        yield g["lr"]


# -------------------------------------------------------------- schedulers
class RampupKeepExpdecayKeepLRScheduler(torch.optim.lr_scheduler.LRScheduler):
    """SpecAugment-style schedule: linear ramp to ``peak_lr`` over ``sr`` steps,
    hold until ``si``, exponential decay to ``floor_lr`` at ``sf``, then hold
    (ref. ``utils.py:189-224``)."""

    def __init__(self, optimizer, peak_lr=0.001, floor_lr=0.00001, sr=1000, si=40000, sf=160000, last_epoch=-1):
        assert peak_lr >= floor_lr and sr <= si <= sf
        self.peak_lr, self.floor_lr, self.sr, self.si, self.sf = peak_lr, floor_lr, sr, si, sf
        self.gamma = math.log(floor_lr / peak_lr) / float(sf - si) if sf > si else 0.0
        self.step_count = 0
        super().__init__(optimizer, last_epoch=last_epoch)

    def lr_at(self, t):
        if t < self.sr:
            return self.peak_lr * float(t) / float(self.sr)
        if t < self.si:
            return self.peak_lr
        if t < self.sf:
            return self.peak_lr * math.exp(self.gamma * float(t - self.si))
        return self.floor_lr

    def get_lr(self):
        return [self.lr_at(self.step_count) for _ in self.base_lrs]

    def step(self, epoch=None):
        for g, lr in zip(self.optimizer.param_groups, self.get_lr()):
            g["lr"] = lr
        self.step_count += 1


class ScheduledSamplingScheduler:
    """Linear ramp of ``model.scheduled_sampling_rate`` (ref. ``utils.py:228-260``)."""

    def __init__(self, model, ramp_start, ramp_stop, initial_rate, final_rate):
        self.model = model
        self.ramp_start, self.ramp_stop = ramp_start, ramp_stop
        self.initial_rate, self.final_rate = initial_rate, final_rate
        self.iter = 0

    def rate_at(self, t):
        if t < self.ramp_start:
            return self.initial_rate
        if t <= self.ramp_stop:
            span = max(self.ramp_stop - self.ramp_start, 1)
            return self.initial_rate + (self.final_rate - self.initial_rate) * ((t - self.ramp_start) / span)
        return self.final_rate

    def step(self):
        r = self.rate_at(self.iter)
        self.model.scheduled_sampling_rate = r
        self.model.scheduled_sampling = (r != 0)
        self.iter += 1

    def state_dict(self):
        return {k: v for k, v in self.__dict__.items() if k != "model"}

    def load_state_dict(self, sd):
        self.__dict__.update(sd)


class NBestTaskScheduler:
    """Multi-task stage scheduler (ref. ``utils.py:263-294``)."""

    def __init__(self, num_tasks, iteration_per_task):
        assert len(num_tasks) == len(iteration_per_task), \
            "Mismatched length {}!={}".format(len(num_tasks), len(iteration_per_task))
        self.iter, self.stagex = 0, 0
        self.num_tasks, self.iteration_per_task = num_tasks, iteration_per_task

    def current_num_tasks(self):
        return self.num_tasks[self.stagex]

    def no_label_updates(self):
        return (self.iter // self.iteration_per_task[-1]) + 1

    def set_iteration_no(self, iter_no):
        self.iter = iter_no

    def step(self):
        local_iter = self.iter % self.iteration_per_task[-1]
        if local_iter == 0:
            self.stagex = 0
        elif local_iter >= self.iteration_per_task[self.stagex]:
            self.stagex += 1
        self.iter += 1


def make_lr_scheduler(annealing_config, optimizer, num_batches=1):
    """step_lr | multi_step_lr | rampup-keep-expdecay-keep | val_loss (ref. ``utils.py:151-186``)."""
    if annealing_config is None:
        return None
    cfg = dict(copy.deepcopy(annealing_config))
    kind = cfg.pop("type")
    interval = cfg.pop("step_interval", "epoch")
    sched = torch.optim.lr_scheduler
    if kind == "step_lr":
        if interval == "epoch":
            cfg["step_size"] = int(num_batches * cfg["step_size"])
        return sched.StepLR(optimizer=optimizer, **cfg)
    # The schema makes ``gamma`` and ``step_size`` REQUIRED for the server's annealing block (``schema.py``), and the
    # reference forwards every key to the scheduler — so ``multi_step_lr`` / ``val_loss`` raise TypeError there on any
    # valid server config.  Here the keys a scheduler does not take are dropped / mapped.
    if kind == "multi_step_lr":
        cfg.pop("step_size", None)
        if interval == "epoch":
            cfg["milestones"] = [int(i * num_batches) for i in cfg["milestones"]]
        return sched.MultiStepLR(optimizer=optimizer, **cfg)
    if kind == "rampup-keep-expdecay-keep":
        cfg.pop("gamma", None)
        cfg.pop("step_size", None)
        return RampupKeepExpdecayKeepLRScheduler(optimizer=optimizer, **cfg)
    if kind == "val_loss":
        gamma, step_size = cfg.pop("gamma", None), cfg.pop("step_size", None)
        cfg.setdefault("factor", gamma if gamma is not None and 0 < gamma < 1 else 0.1)
        cfg.setdefault("patience", int(step_size) if step_size is not None else 10)
        return sched.ReduceLROnPlateau(optimizer, **cfg)
    raise ValueError("{} LR scheduler not supported".format(kind))


# ------------------------------------------------------------------- math
def softmax(X, theta=1.0, axis=None):
    y = np.atleast_2d(np.asarray(X, dtype=np.float64))
    if axis is None:
        axis = next((i for i, n in enumerate(y.shape) if n > 1), 0)
    y = y * float(theta)
    y = np.exp(y - np.max(y, axis=axis, keepdims=True))
    p = y / np.sum(y, axis=axis, keepdims=True)
    return p.flatten() if np.ndim(X) == 1 else p


class AverageMeter:
    """Micro/macro running averages of a ratio metric (ref. ``utils.py:117-148``)."""

    def __init__(self, metric_name):
        self.metric_name = metric_name
        self.reset()

    def reset(self):
        self.numerators, self.denominators = [], []

    def add(self, top, bottom):
        self.numerators.append(top)
        self.denominators.append(bottom)

    def get_macro_average(self):
        return self.get_average([float(n) / d for n, d in zip(self.numerators, self.denominators)])

    def get_micro_average(self):
        return float(sum(self.numerators)) / sum(self.denominators)

    @staticmethod
    def get_average(l):
        return sum(l) / float(len(l))

    def display_results(self, loglevel=logging.INFO):
        print_rank("{} Macro average: {}".format(self.metric_name, self.get_macro_average()), loglevel)
        print_rank("{} Micro average: {}".format(self.metric_name, self.get_micro_average()), loglevel)


# -------------------------------------------------------- gradient helpers
def flatten_grads_model(learner) -> np.ndarray:
    return torch.cat([p.grad.detach().reshape(-1) for p in learner.parameters()]).cpu().numpy()


def flatten_grads_array(param_array) -> np.ndarray:
    return np.stack([torch.cat([w.detach().reshape(-1) for w in ws]).cpu().numpy() for ws in param_array])


def dist_weights_to_model(weights, parameters):
    off = 0
    for p in parameters:
        n = p.numel()
        p.data.copy_(torch.as_tensor(weights[off:off + n]).reshape(p.shape))
        off += n


def dist_params_to_model(grads, model):
    off = 0
    for p in model:
        n = p.numel()
        g = torch.as_tensor(grads[off:off + n]).reshape(p.shape).to(p)
        p.grad = g if p.grad is None else p.grad + g
        off += n


def reshape_params_to_model(grads, model):
    out, off = [], 0
    for p in model:
        n = p.numel()
        out.append(torch.as_tensor(grads[off:off + n]).reshape(p.shape).to(p))
        off += n
    return out


def scrub_empty_clients(data_strct):
    """Drop users with zero samples (ref. ``utils.py:563-582``)."""
    from ..core.config import ConfigNode
    has_labels = "user_data_label" in data_strct
    out = {"users": [], "user_data": {}, "num_samples": []}
    if has_labels:
        out["user_data_label"] = {}
    for i, u in enumerate(data_strct["users"]):
        if data_strct["num_samples"][i] > 0:
            out["users"].append(u)
            out["user_data"][u] = data_strct["user_data"][u]
            out["num_samples"].append(data_strct["num_samples"][i])
            if has_labels:
                out["user_data_label"][u] = data_strct["user_data_label"][u]
    return ConfigNode(out)


def compute_grad_cosines(grads, model_grad):
    """cos(g_k, G) per client (ref. ``utils.py:585-595``).  CUDA inputs: one fused kernel per client
    (``ops.misc_ops.cosine_stats``: dot and both norms in one pass, nothing copied to the host); the whole list is
    read back once."""
    def flat(ts):
        ts = list(ts) if not torch.is_tensor(ts) else [ts]
        return ts[0].detach().reshape(-1).float() if len(ts) == 1 else torch.cat([t.detach().reshape(-1).float() for t in ts])

    G = flat(model_grad)
    if G.is_cuda:
        from ..ops import misc_ops
        cos = [misc_ops.cosine(flat(g).to(G.device), G) for g in grads]
        return [float(c) for c in torch.stack(cos).cpu()] if cos else []
    G = G.cpu()
    Gn = G.norm()
    out = []
    for g in grads:
        f = flat(g).cpu()
        fn = f.norm()
        out.append(float(torch.dot(f, G) / (fn * Gn)) if fn > 0 and Gn > 0 else 0)
    return out


def convex_inference(model_global, model_personal, alpha):
    """Accuracy of α·personal + (1−α)·global probabilities (ref. ``utils.py:598-603``)."""
    targets = torch.as_tensor(np.asarray(model_global["labels"]))
    probs = alpha * np.asarray(model_personal["probabilities"]) + (1 - alpha) * np.asarray(model_global["probabilities"])
    pred = torch.argmax(torch.as_tensor(probs), dim=1)
    return torch.mean((pred == targets).float()).item()


def alpha_update(model_global, model_personal, alpha, eta):
    """One SGD step on the personalization mixing weight (ref. ``utils.py:605-617``).

    K25 in SURVEY §2.4: a single fused dot-product over the flattened models
    instead of one ``dot`` per tensor."""
    from ..ops import misc_ops
    from ..parallel.arena import module_arena
    ag, ap = module_arena(model_global), module_arena(model_personal)
    if (ag is not None and ap is not None and ag[1] is not None and ap[1] is not None
            and ag[0].flat.numel() == ap[0].flat.numel()):
        # both models live in flat arenas: the four operands already ARE flat vectors (padding is zero in all of them)
        grad_alpha = misc_ops.alpha_dot(ap[0].flat, ag[0].flat, ap[1].flat, ag[1].flat, float(alpha))[0] + 0.02 * alpha
    else:
        lp = list(model_global.parameters())
        pp = list(model_personal.parameters())
        zero = lambda t: t.grad if t.grad is not None else torch.zeros_like(t)
        dif = torch.cat([(p.data - l.data).reshape(-1) for l, p in zip(lp, pp)])
        gp = torch.cat([zero(p).reshape(-1) for p in pp])
        gl = torch.cat([zero(l).reshape(-1) for l in lp])
        grad_alpha = misc_ops.alpha_dot(dif, torch.zeros_like(dif), gp, gl, float(alpha))[0] + 0.02 * alpha
    alpha_n = float(np.clip((alpha - eta * grad_alpha).item(), 0.0001, 0.9999))
    return alpha_n if np.isfinite(alpha_n) else 0.75


def get_label_VAT(local_logits, server_logits, thre, comp):
    """Pick pseudo-labels from whichever of (local, server) is more confident
    (FedLabels; ref. ``utils.py:620-678``).  Vectorised: no per-sample Python loop."""
    if comp == "var":
        lc, sc = torch.var(local_logits, dim=1), torch.var(server_logits, dim=1)
    elif comp == "ent":
        def inv_entropy(p):
            p = p / p.sum(dim=1, keepdim=True)
            ent = -(torch.where(p > 0, p * torch.log(p), torch.zeros_like(p))).sum(dim=1) + 0.00001
            return 1.0 / ent
        lc, sc = inv_entropy(local_logits), inv_entropy(server_logits)
    else:
        raise ValueError("comp must be 'var' or 'ent'")
    use_local = (lc >= sc) & (local_logits.max(dim=1).values > thre)
    use_server = (lc < sc) & (server_logits.max(dim=1).values > thre)
    keep = use_local | use_server
    idx = torch.nonzero(keep).flatten()
    if idx.numel() == 0:
        return [], [], [], 0
    labels = torch.where(use_local, local_logits.argmax(dim=1), server_logits.argmax(dim=1))[idx]
    var = torch.where(use_local, sc / lc, lc / sc)[idx]
    n_server, n_local = int(use_server.sum()), int(use_local.sum())
    ratio = n_server / (n_server + n_local)
    return labels, idx.tolist(), (var if comp == "var" else []), ratio


# ------------------------------------------------- n-best jsonl (ASR legacy)
def write_nbest_jsonl(uttid2jsonl, uttid2hypos, uttid2scores, outputpath, nbest, orgpath="", newpath=""):
    """Dump a json-list with (n-best) hypotheses (ref. ``utils.py:362-397``)."""
    out = []
    for uttid, rec in uttid2jsonl.items():
        if uttid not in uttid2hypos:
            print_rank("Missing utterance {} in results".format(uttid))
            continue
        hypos = uttid2hypos[uttid]
        if nbest > 1:
            w = _nbest_weights(uttid, uttid2scores, nbest)
            for n in range(min(nbest, len(hypos))):
                new = copy.deepcopy(rec)
                new.update(id="{}-{}".format(uttid, n), text=" ".join(hypos[n]), loss_weight=float(w[n]))
                out.append(new)
        else:
            new = copy.deepcopy(rec)
            new.update(id=uttid, text=" ".join(hypos[0]))
            out.append(new)
    _dump_jsonl(out, outputpath, orgpath, newpath)
    return True


def _nbest_weights(uttid, uttid2scores, nbest):
    if uttid not in uttid2scores:
        return np.ones(nbest) / nbest
    w = np.asarray(uttid2scores[uttid], dtype=np.float64)
    if len(w) < nbest:
        w = np.concatenate([w, np.full(nbest - len(w), w[0])])
    return softmax(w[:nbest])


def _dump_jsonl(records, outputpath, orgpath, newpath):
    with open(outputpath, "w") as f:
        for r in records:
            if "wav" in r:
                r["wav"] = r["wav"].replace(orgpath, newpath)
            f.write(json.dumps(r) + "\n")


def write_multitask_jsonl(uttid2jsonl, uttid2hypos, uttid2scores, outputpath, nbest, orgpath="", newpath=""):
    """Multi-task variant: one record per utterance with ``task_weights`` and ``subtextl``
    (ref. ``utils.py:399-446``)."""
    if nbest == 1:
        return write_nbest_jsonl(uttid2jsonl, uttid2hypos, uttid2scores, outputpath, nbest, orgpath, newpath)
    out = []
    for uttid, rec in uttid2jsonl.items():
        if uttid not in uttid2hypos:
            print_rank("Missing utterance {} in results".format(uttid))
            continue
        hypos = uttid2hypos[uttid]
        w = _nbest_weights(uttid, uttid2scores, nbest)
        rec["task_weights"] = [float(x) for x in w]
        rec["text"] = " ".join(hypos[0])
        rec["subtextl"] = [" ".join(hypos[n] if n < len(hypos) else hypos[0]) for n in range(1, nbest)]
        if rec["text"] == "" and all(s == "" for s in rec["subtextl"]):
            print_rank("Skip {}: Invalid result".format(uttid))
            continue
        out.append(rec)
    _dump_jsonl(out, outputpath, orgpath, newpath)
    return True


def load_eval_result_jsonl(resultjsonl, uttid2hypos=None, uttid2scores=None, dumpfp=None, dump_msg="RESULT: "):
    """Read back an evaluator json-list (ref. ``utils.py:448-483``)."""
    uttid2hypos = OrderedDict() if uttid2hypos is None else uttid2hypos
    uttid2scores = OrderedDict() if uttid2scores is None else uttid2scores
    best = oracle = length = 0
    with open(resultjsonl) as f:
        for line in f:
            e = json.loads(line.strip())
            if "hypothesis" in e:
                k = next(iter(e["hypothesis"]))
                uttid2hypos[e["utt_id"]] = e["hypothesis"][k]
                if "nbest_model_scores" in e:
                    uttid2scores[e["utt_id"]] = np.array(e["nbest_model_scores"][k])
            else:
                if dumpfp is not None:
                    dumpfp.write("{}{}\n".format(dump_msg, line.strip()))
                k = next(iter(e["wer-"]))
                rec = e["wer-"][k]
                best += rec["best_wer"] * rec["total_length"]
                oracle += rec["oracle_wer"] * rec["total_length"]
                length += rec["total_length"]
    return uttid2hypos, uttid2scores, best, oracle, length
