"""Host-side mirror of `brightness_matcher.BrightnessMatcher` (reference: brightness_matcher.py:14-73)."""
from __future__ import annotations

from . import ops


class BrightnessMatcher:
    """match scene brightness to patch: `BrightnessMatcher()((src, tgt)) -> src'` on the GPU."""

    def __init__(self, *args, name=None, **kwargs):
        self.name = name

    def __call__(self, inputs, **kwargs):
        return self.call(inputs, **kwargs)

    def call(self, inputs, **kwargs):
        src, tgt = inputs
        return ops.brightness_match(src, tgt)
