from .LinearOperations import *    # noqa: F401,F403
from .Statistics import *          # noqa: F401,F403
from .QuasiNewtonTools import *    # noqa: F401,F403
from .BFGS import *                # noqa: F401,F403
from .DIIS import *                # noqa: F401,F403
