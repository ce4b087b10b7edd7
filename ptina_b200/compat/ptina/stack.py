from ptina.common import *  # noqa: F401,F403
# the reference's per-thread global-memory stack (stack.py:10-123) has no counterpart: traversal stacks live in registers / local memory
