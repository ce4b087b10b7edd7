from ptina.engine import *  # noqa: F401,F403
from ptina.sampling import *  # noqa: F401,F403
from ptina.sampling.sobol import *  # noqa: F401,F403
from ptina_b200.engine.path import *  # noqa: F401,F403
