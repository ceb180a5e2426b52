from .utils import *  # noqa: F401,F403
from .utils import (env_rank, env_local_rank, env_world_size, default_device)  # noqa: F401
from .optimizers import AdamW, LAMB, LarsSGD, LarsSGDV1, LARSWrapper  # noqa: F401
