from .SimpleMD import *      # noqa: F401,F403
from .PeriodicMD import *    # noqa: F401,F403
from .PeriodicMC import PeriodicMonteCarlo   # noqa: F401
from .Opt import *           # noqa: F401,F403
from .OptPeriodic import *   # noqa: F401,F403
from .Neb import *           # noqa: F401,F403
from .DeviceMD import DevicePeriodicVelocityVerlet, DeviceVelocityVerlet   # noqa: F401  (B200 extension: SURVEY 8f N1)
