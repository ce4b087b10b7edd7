"""ptina_b200 -- B200-native (sm_100a) implementation of PTina's per-pixel path-tracing hot path behind
PTina's own Python engine API.  See DESIGN.md."""
__version__ = '0.1.0'
