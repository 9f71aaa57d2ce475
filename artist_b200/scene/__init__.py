from .light_source import LightSource
from .light_source_array import LightSourceArray
from .rays import Rays
from .sun import Sun

__all__ = ["LightSource", "LightSourceArray", "Rays", "Sun"]
