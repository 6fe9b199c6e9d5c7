from .SimpleMD import *      # noqa: F401,F403
from .PeriodicMD import *    # noqa: F401,F403
from .Opt import *           # noqa: F401,F403
from .OptPeriodic import *   # noqa: F401,F403
from .Neb import *           # noqa: F401,F403
