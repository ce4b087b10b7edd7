from ptina.camera import *  # noqa: F401,F403
from ptina.model import *  # noqa: F401,F403
from ptina.light import *  # noqa: F401,F403
from ptina.light.world import *  # noqa: F401,F403
from ptina.filmtable import *  # noqa: F401,F403
from ptina.mtllib import *  # noqa: F401,F403
from ptina.stack import *  # noqa: F401,F403
from ptina.tree import *  # noqa: F401,F403
