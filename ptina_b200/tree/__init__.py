"""BVH for ray-tracing acceleration (reference: ptina/tree/__init__.py:5)."""
from .lbvh import BVHTree, LinearBVH  # noqa: F401
