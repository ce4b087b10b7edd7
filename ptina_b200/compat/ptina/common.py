import taichi as ti  # noqa: F401
from ptina_b200.common import *  # noqa: F401,F403
from ptina_b200.common import np, eps, inf, Singleton  # noqa: F401
