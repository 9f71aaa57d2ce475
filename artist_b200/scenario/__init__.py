from .scenario import Scenario
from .synthetic import build_synthetic_scenario, synthetic_field_tensors

__all__ = ["Scenario", "build_synthetic_scenario", "synthetic_field_tensors"]
