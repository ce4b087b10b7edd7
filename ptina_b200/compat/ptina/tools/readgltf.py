from ptina_b200.tools.readgltf import *  # noqa: F401,F403
