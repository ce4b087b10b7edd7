from ptina_b200.tools.readobj import *  # noqa: F401,F403
