from .TMIPIinterface import *   # noqa: F401,F403
