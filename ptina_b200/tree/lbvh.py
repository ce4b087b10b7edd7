"""Linear BVH (reference: ptina/tree/lbvh.py:45-358).  `BVHTree().build()` runs the whole build on the device
(Morton codes, radix sort, Karras hierarchy with the reference's quirks, AABB fit) and returns nothing, like the
reference; it raises RuntimeError('AABB step never stop! hierarchy corrupted?') in the same situation."""
from ..common import Singleton
from .. import _native


class LinearBVH:
    def __init__(self, n=2**22):
        self.capacity = n
        self.info = None

    def build(self):
        try:
            self.info = _native.context().build_tree()
        except _native.NativeError as e:
            raise RuntimeError(str(e)) from None

    def export(self):
        """The reference's own arrays: mc, id, child, leaf, bmin, bmax (lbvh.py:50-59), as NumPy."""
        return _native.context().export_tree()


class BVHTree(LinearBVH, metaclass=Singleton):
    pass
