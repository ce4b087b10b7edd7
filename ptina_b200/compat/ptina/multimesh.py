from ptina.common import *  # noqa: F401,F403
from ptina_b200.multimesh import *  # noqa: F401,F403
