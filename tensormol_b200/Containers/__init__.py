from .Mol import *          # noqa: F401,F403
from .Sets import *         # noqa: F401,F403
from .DigestMol import *    # noqa: F401,F403
from .TensorMolData import *  # noqa: F401,F403
