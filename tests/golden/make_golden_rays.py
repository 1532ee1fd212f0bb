#!/usr/bin/env python
"""Golden fixture for ray generation (the caller on the input side of the hot path) from the LIVE reference.

    python tests/golden/make_golden_rays.py      # rewrites tests/golden/rays.npz

Six nuScenes-shaped pinhole cameras (1600 x 900, fx = fy = 1266 — 809 for the back camera —, yaw offsets 0, +-55, +-110,
180 degrees) at a few ego poses; `Cameras.generate_rays` (cameras/cameras.py:497-880) through the same
indexing `RayGenerator.forward` uses (model_components/ray_generators.py:43-61), for random pixels plus the four image
corners.  Stores the camera parameters, the ray indices and origins / directions / pixel_area / directions_norm.
"""
import math
import os
import sys

sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (installs the import shims and sys.path for the reference)
import torch  # noqa: E402
from nerfstudio.cameras.cameras import Cameras, CameraType  # noqa: E402


def yaw_pose(yaw_deg, pos):
    """camera-to-world of an OpenGL-convention camera (looks down -z, +y up) rotated about the world z axis."""
    a = math.radians(yaw_deg)
    fwd = torch.tensor([math.cos(a), math.sin(a), 0.0])
    up = torch.tensor([0.0, 0.0, 1.0])
    right = torch.linalg.cross(fwd, up)
    rot = torch.stack([right, up, -fwd], dim=1)          # columns = camera x, y, z axes in the world
    return torch.cat([rot, pos[:, None]], dim=1)


def main():
    g = torch.Generator().manual_seed(11)
    yaws = [0.0, 55.0, -55.0, 110.0, -110.0, 180.0]
    c2w, fx = [], []
    for frame in range(3):
        pos = torch.tensor([frame * 2.0 - 1.0, frame * 0.7, 1.5]) * 0.05
        for k, yw in enumerate(yaws):
            c2w.append(yaw_pose(yw + 7.0 * frame, pos))
            fx.append(809.0 if k == 5 else 1266.0)
    c2w, fx = torch.stack(c2w), torch.tensor(fx)
    C = c2w.shape[0]
    cams = Cameras(camera_to_worlds=c2w, fx=fx, fy=fx.clone(), cx=800.0, cy=450.0, width=1600, height=900,
                   camera_type=CameraType.PERSPECTIVE)
    n = 4096
    idx = torch.stack([torch.randint(0, C, (n,), generator=g), torch.randint(0, 900, (n,), generator=g),
                       torch.randint(0, 1600, (n,), generator=g)], dim=1)
    idx[:4] = torch.tensor([[0, 0, 0], [5, 899, 1599], [7, 0, 1599], [17, 899, 0]])
    coords = cams.get_image_coords()[idx[:, 1], idx[:, 2]]                      # RayGenerator.forward, :53
    rb = cams.generate_rays(camera_indices=idx[:, :1], coords=coords)
    MG.save("rays.npz", {"c2w": c2w, "fx": fx, "fy": fx, "cx": torch.full((C,), 800.0), "cy": torch.full((C,), 450.0),
                         "ray_indices": idx, "origins": rb.origins, "directions": rb.directions,
                         "pixel_area": rb.pixel_area, "directions_norm": rb.metadata["directions_norm"]})


if __name__ == "__main__":
    main()
