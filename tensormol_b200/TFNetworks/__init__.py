from .TFMolManage import *   # noqa: F401,F403
