"""lucille_b200 -- B200-native ray-intersection backend behind lucille's ri_accel_* / ri_raytrace() boundary."""
__version__ = "0.1.0"
