from numpy import *          # noqa: F401,F403
from numpy import linalg, random  # noqa: F401
