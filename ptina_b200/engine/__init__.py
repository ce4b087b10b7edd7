"""Integrators (reference: ptina/engine/).  Every engine's `render()` adds one sample per pixel to the film, exactly
like the reference; `render(nsamples)` batches several such calls into one wavefront submission."""
from .path import PathEngine
from .brute import BruteEngine
from .preview import PreviewEngine
from .mltpath import MLTPathEngine

ENGINES = (PathEngine, BruteEngine, PreviewEngine, MLTPathEngine)
