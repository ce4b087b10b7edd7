from ptina.sampling import *  # noqa: F401,F403
from ptina_b200.sampling.sobol import *  # noqa: F401,F403
