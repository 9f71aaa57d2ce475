"""Tensor-axis / row names used by the host classes (subset of ``artist/util/indices.py`` semantics)."""
# homogeneous ENU components
e, n, u, w = 0, 1, 2, 3
# target-area types inside SolarTower.target_areas
planar_target_areas, cylindrical_target_areas = 0, 1
# bitmap_resolution tensor = [E, U]
unbatched_bitmap_e, unbatched_bitmap_u = 0, 1
bitmap_resolution = 256
# actuator non-optimizable rows [N,7,2]
actuator_type, actuator_clockwise_movement, actuator_min_motor_position, actuator_max_motor_position = 0, 1, 2, 3
actuator_increment, actuator_offset, actuator_pivot_radius = 4, 5, 6
# actuator optimizable rows [N,2,2]
actuator_initial_angle, actuator_initial_stroke_length = 0, 1
# kinematics deviation columns
first_joint_tilt_n, first_joint_tilt_u, second_joint_tilt_e, second_joint_tilt_n = 0, 1, 2, 3
concentrator_translation_n = 7
