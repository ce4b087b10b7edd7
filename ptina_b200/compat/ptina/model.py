from ptina.common import *  # noqa: F401,F403
from ptina_b200.model import *  # noqa: F401,F403
