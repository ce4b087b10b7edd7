from ptina_b200.tools.matrix import *  # noqa: F401,F403
