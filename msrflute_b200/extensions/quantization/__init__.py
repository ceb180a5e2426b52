from .quant import *  # noqa: F401,F403
