from ptina.common import *  # noqa: F401,F403
from ptina_b200.light.world import *  # noqa: F401,F403
