"""presight_b200 — B200-native (sm_100a) implementation of PreSight's city-scale NeRF inner loop behind the
reference's nerfstudio operator surface.  See DESIGN.md and include/presight_b200.h."""
__version__ = "0.1.0"
