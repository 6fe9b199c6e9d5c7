from .Neighbors import *   # noqa: F401,F403
from .Periodic import *    # noqa: F401,F403
