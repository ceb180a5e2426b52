"""Strategy interface (ref. ``core/strategies/base.py:20-56``).

A strategy has a *client* face (turn a finished local training run into a
payload ``{'weight', 'gradients'}``) and a *server* face (consume payloads one
by one, then combine them into a model update).

Two B200 additions, both optional:

``client_weight(trainer)``       just the aggregation weight — used by the fused
                                 device path where the weighted pseudo-gradient is
                                 accumulated by a kernel and never materialised.
``needs_individual_payloads``    True when the server must see every client's
                                 gradient separately (stale gradients, cosine
                                 dumps, RL) — this disables worker-side
                                 pre-reduction, mirroring the reference's own
                                 ``fast_aggregation`` restrictions (``dga.py:71-76``).
"""


class BaseStrategy:
    def __init__(self, mode, config, model_path=None):
        if mode not in ("client", "server"):
            raise ValueError("mode in strategy must be either `client` or `server`")
        self.mode, self.config, self.model_path = mode, config, model_path

    # client side
    def client_weight(self, trainer):
        raise NotImplementedError

    def generate_client_payload(self, trainer):
        raise NotImplementedError

    # server side
    def process_individual_payload(self, worker_trainer, payload):
        raise NotImplementedError

    def combine_payloads(self, worker_trainer, curr_iter, num_clients_curr_iter, total_clients, client_stats,
                         logger=None):
        raise NotImplementedError

    @property
    def needs_individual_payloads(self):
        return False

    def _require(self, mode):
        if self.mode != mode:
            raise RuntimeError("this method can only be invoked by the {}".format(mode))

    def __repr__(self):
        return "{}(mode={})".format(type(self).__name__, self.mode)
