"""Named map configurations.

``SETTINGS`` restates the four compile-time presets of the reference
(include/settings/settings.h:32-119); ``BASELINE`` holds the five workloads of
/root/repo/BASELINE.json as SURVEY.md §8(d) pins them (grid bits, camera, parameters, speed).
"""
from __future__ import annotations

from dataclasses import dataclass, replace

from . import _capi


@dataclass(frozen=True)
class MapConfig:
    name: str
    voxel_bits: tuple          # (x, y, z)
    slot_bits: int
    voxel_size: float
    fx: float
    fy: float
    cx: float
    cy: float
    image_width: int
    image_height: int
    boost_mode: bool
    depth_min: float
    depth_max: float
    consider_instance: bool
    zed2_filters: bool
    window_half: int = 0       # 0 = reference default (3 boost / 5)
    # runtime parameters (cfg/options_*.yaml of the reference)
    detection_probability: float = 0.8
    noise_number: float = 0.2
    nb_ptc_num_per_point: int = 1
    occupancy_threshold: float = 0.15
    forgetting_rate: float = 1.0
    max_forget_count: int = 5
    id_transition_probability: float = 0.5
    depth_noise_first_order: float = 0.02
    depth_noise_zero_order: float = 0.3
    consider_depth_noise: bool = True
    use_independent_filter: bool = False
    zero_center: bool = False
    # synthetic trajectory
    speed: float = 0.3         # metres per frame along +z
    seed: int = 20240911

    @property
    def internal_size(self):
        """(rows, cols) of the internal image: settings.h:128-143."""
        if self.boost_mode:
            return int(0.5 * self.image_height), int(0.5 * self.image_width)
        return self.image_height, self.image_width

    @property
    def n_voxels(self):
        return 1 << sum(self.voxel_bits)

    @property
    def n_slots(self):
        return self.n_voxels << self.slot_bits

    def c_config(self, shard_rank=0, shard_count=1, device=0, max_visible=0) -> _capi.Config:
        c = _capi.Config()
        c.voxel_bits_x, c.voxel_bits_y, c.voxel_bits_z = self.voxel_bits
        c.slot_bits = self.slot_bits
        c.voxel_size = self.voxel_size
        c.fx, c.fy, c.cx, c.cy = self.fx, self.fy, self.cx, self.cy
        c.image_width, c.image_height = self.image_width, self.image_height
        c.boost_mode = int(self.boost_mode)
        c.depth_min, c.depth_max = self.depth_min, self.depth_max
        c.consider_instance = int(self.consider_instance)
        c.zed2_filters = int(self.zed2_filters)
        c.window_half = self.window_half
        c.shard_rank, c.shard_count, c.device = shard_rank, shard_count, device
        c.max_visible = max_visible
        return c

    def c_params(self) -> _capi.Params:
        p = _capi.Params()
        p.detection_probability = self.detection_probability
        p.noise_number = self.noise_number
        p.nb_ptc_num_per_point = self.nb_ptc_num_per_point
        p.occupancy_threshold = self.occupancy_threshold
        p.forgetting_rate = self.forgetting_rate
        p.max_forget_count = self.max_forget_count
        p.id_transition_probability = self.id_transition_probability
        p.depth_noise_first_order = self.depth_noise_first_order
        p.depth_noise_zero_order = self.depth_noise_zero_order
        p.consider_depth_noise = int(self.consider_depth_noise)
        p.use_independent_filter = int(self.use_independent_filter)
        p.zero_center = int(self.zero_center)
        return p


_ZED2_CAM = dict(fx=527.8191528320312, fy=527.8191528320312, cx=633.9357299804688, cy=366.3338623046875,
                 image_width=1280, image_height=720, boost_mode=True, depth_min=0.3, depth_max=15.0,
                 consider_instance=True, zed2_filters=True)
# cfg/options_zed2.yaml
_ZED2_PRM = dict(detection_probability=0.8, noise_number=0.2, nb_ptc_num_per_point=1, occupancy_threshold=0.15,
                 forgetting_rate=1.0, max_forget_count=5, id_transition_probability=0.5,
                 depth_noise_first_order=0.02, depth_noise_zero_order=0.3, consider_depth_noise=True,
                 use_independent_filter=False, zero_center=True)
# cfg/options_virtual_kitti2.yaml
_VK2_PRM = dict(detection_probability=0.98, noise_number=0.001, nb_ptc_num_per_point=1, occupancy_threshold=0.5,
                forgetting_rate=1.0, max_forget_count=3, id_transition_probability=0.2,
                depth_noise_first_order=0.01, depth_noise_zero_order=0.2, consider_depth_noise=True,
                use_independent_filter=False, zero_center=True)
# cfg/options_kitti360.yaml
_K360_PRM = dict(detection_probability=0.6, noise_number=0.6, nb_ptc_num_per_point=1, occupancy_threshold=0.1,
                 forgetting_rate=0.5, max_forget_count=3, id_transition_probability=0.2,
                 depth_noise_first_order=0.01, depth_noise_zero_order=0.1, consider_depth_noise=True,
                 use_independent_filter=True, zero_center=False)

# include/settings/settings.h:32-119
SETTINGS = {
    0: MapConfig("setting0_kitti360", (8, 8, 8), 3, 0.15, 552.554261, 552.554261, 682.049453, 238.769549, 1408, 376,
                 False, 0.3, 30.0, False, False, **_K360_PRM),
    1: MapConfig("setting1_coda", (8, 8, 7), 2, 0.15, 569.8286, 565.4818, 439.2660, 360.5810, 960, 540,
                 False, 0.3, 10.0, True, False, **_VK2_PRM),
    2: MapConfig("setting2_vkitti2", (8, 7, 8), 3, 0.2, 725.0087, 725.0087, 620.5, 187.0, 1242, 375,
                 False, 0.3, 30.0, True, False, **_VK2_PRM),
    3: MapConfig(name="setting3_zed2", voxel_bits=(7, 5, 7), slot_bits=2, voxel_size=0.15, **_ZED2_CAM, **_ZED2_PRM),
}

# /root/repo/BASELINE.json configs[0..4] == SURVEY.md §8 A..E
BASELINE = {
    "A": MapConfig(name="A_static_64x64x32_4ppv_640x480", voxel_bits=(6, 5, 6), slot_bits=2, voxel_size=0.15,
                   fx=525.0, fy=525.0, cx=319.5, cy=239.5, image_width=640, image_height=480, boost_mode=False,
                   depth_min=0.3, depth_max=10.0, consider_instance=False, zed2_filters=False,
                   speed=0.3, seed=20240911 + 0, **_K360_PRM),
    "B": MapConfig(name="B_zed2_256x256x64_8ppv_1280x720", voxel_bits=(8, 6, 8), slot_bits=3, voxel_size=0.15,
                   speed=0.3, seed=20240911 + 1, **_ZED2_CAM, **_ZED2_PRM),
    "C": MapConfig(name="C_vkitti2_256x256x64_8ppv_1242x375", voxel_bits=(8, 6, 8), slot_bits=3, voxel_size=0.2,
                   fx=725.0087, fy=725.0087, cx=620.5, cy=187.0, image_width=1242, image_height=375,
                   boost_mode=False, depth_min=0.3, depth_max=30.0, consider_instance=True, zed2_filters=False,
                   speed=0.3, seed=20240911 + 2, **_VK2_PRM),
    "D": MapConfig(name="D_512x512x128_16ppv_zed2", voxel_bits=(9, 7, 9), slot_bits=4, voxel_size=0.15,
                   speed=2.0, seed=20240911 + 3, **_ZED2_CAM, **_ZED2_PRM),
    "E": MapConfig(name="E_1024x1024x128_16ppv_zed2", voxel_bits=(10, 7, 10), slot_bits=4, voxel_size=0.15,
                   speed=1.0, seed=20240911 + 4, **_ZED2_CAM, **_ZED2_PRM),
}


def tiny(base: MapConfig = None, **kw) -> MapConfig:
    """A seconds-scale variant for unit tests: small grid, small image."""
    base = base or BASELINE["B"]
    d = dict(name="tiny_" + base.name, voxel_bits=(6, 5, 6), slot_bits=base.slot_bits, image_width=160, image_height=96,
             fx=base.fx * 160 / base.image_width, fy=base.fy * 96 / base.image_height,
             cx=base.cx * 160 / base.image_width, cy=base.cy * 96 / base.image_height)
    d.update(kw)
    return replace(base, **d)
