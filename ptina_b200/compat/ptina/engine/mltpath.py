from ptina.engine.path import *  # noqa: F401,F403
from ptina_b200.engine.mltpath import *  # noqa: F401,F403
