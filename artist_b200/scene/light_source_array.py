from collections.abc import Sequence

from .light_source import LightSource


class LightSourceArray:
    """List of light sources (``artist/scene/light_source_array.py:16-44``)."""

    def __init__(self, light_source_list: Sequence[LightSource]) -> None:
        self.light_source_list = light_source_list
