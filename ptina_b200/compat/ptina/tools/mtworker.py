from ptina_b200.tools.mtworker import *  # noqa: F401,F403
