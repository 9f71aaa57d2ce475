from .bitmap import crop_flux_distributions_around_center, get_center_of_mass, trapezoid_distribution

__all__ = ["crop_flux_distributions_around_center", "get_center_of_mass", "trapezoid_distribution"]
