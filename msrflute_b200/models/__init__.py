"""Model zoo + factory.

``make_model`` keeps the reference's plug-in convention
(``experiments/__init__.py:8-43``): ``model_config.model_folder`` is the path of
a python file and ``model_config.model_type`` the class inside it, instantiated
as ``model_type(model_config)``; optional ``weight_init: xavier_normal``.
"""
import torch

from ..utils import print_rank, to_device
from ..utils.dataloaders_utils import load_source, resolve_path


def make_model(model_config, dataloader_type=None, input_dim=-1, output_dim=-1, vocab_size=None, device=True):
    folder, klass = str(model_config["model_folder"]), model_config["model_type"]
    try:
        mod = load_source(klass, resolve_path(folder))
        model_type = getattr(mod, klass)
    except (FileNotFoundError, AttributeError) as e:
        raise ValueError("{} model not found, make sure to indicate the model path in the .yaml file ({})"
                         .format(klass, e))
    model = model_type(model_config)
    init = model_config.get("weight_init", "default")
    if init == "xavier_normal":
        for p in model.parameters():
            if p.dim() > 1:
                torch.nn.init.xavier_normal_(p.data)
            elif p.dim() == 1:
                p.data.zero_()
        for m in model.modules():
            if isinstance(m, (torch.nn.Embedding, torch.nn.LayerNorm, torch.nn.BatchNorm2d)):
                m.reset_parameters()
    elif init != "default":
        raise ValueError("{} not supported".format(init))
    return to_device(model) if device else model
