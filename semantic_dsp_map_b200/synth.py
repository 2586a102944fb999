"""Seeded synthetic frames of the shape the reference consumes.

A frame = depth (float32 metres), one "static" class mask, one {0,255} mask per tracked box with its 3-D key points,
the camera pose, and the per-object rigid moves the (host-side, out-of-scope) object layer would have produced.
The scene is ray-cast analytically: ground plane y = +1.5 m (camera y points down), side walls at x = +-wall_x,
a back wall at a fixed camera depth, and ``n_boxes`` car-sized boxes, half of which translate 0.1 m/frame along x.
Everything is a pure function of (config, frame index), so the oracle and the engine see identical inputs.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .configs import MapConfig
from .map import MaskKpts

# static-mask pixel values; the label id is value+1 in the ZED2 table (pointcloud_tools.h:137; cfg/object_info_zed2.csv)
V_ROAD, V_BUILDING, V_WALL, V_TERRAIN, V_VEHICLE = 0, 2, 3, 9, 13
BOX_SIZE = np.array([1.5, 1.5, 4.0])   # x (width), y (height), z (length)
BOX_SPEED_X = 0.1


@dataclass
class Frame:
    index: int
    depth: np.ndarray
    masks: List[MaskKpts]
    position: np.ndarray       # float64[3]
    orientation: np.ndarray    # float64[4] w,x,y,z
    moves: list                # [(track_id, float32 4x4)]
    remove_tracks: list


def make_random_table(seed: int, n: int = 1_000_000) -> np.ndarray:
    """The reference pre-draws 10^6 N(0, 0.05) values (utils/basic_algorithms.h:394-401); injected for determinism."""
    return np.random.default_rng(seed).normal(0.0, 0.05, n).astype(np.float32)


class Scene:
    def __init__(self, cfg: MapConfig, n_boxes: int = 8, instances: bool = None, noise: bool = True):
        self.cfg = cfg
        self.n_boxes = n_boxes
        self.instances = cfg.consider_instance if instances is None else instances
        self.noise = noise
        half = [(1 << (b - 1)) * cfg.voxel_size for b in cfg.voxel_bits]
        self.wall_x = min(6.0, 0.6 * half[0])
        self.back_depth = 0.8 * min(cfg.depth_max, half[2])
        rng = np.random.default_rng(cfg.seed)
        # boxes are laid out along the trajectory (about two in view at a time), left and right of it, resting on
        # the ground; they shrink with the visible range so that small test maps see them too
        self.box_size = BOX_SIZE * min(1.0, self.back_depth / 12.0)
        self.box_z0 = 0.5 * self.back_depth + 0.45 * self.back_depth * (np.arange(n_boxes) + rng.uniform(0.0, 0.5, n_boxes))
        self.box_x0 = rng.uniform(-0.7 * self.wall_x, 0.7 * self.wall_x - self.box_size[0], n_boxes)
        self.box_moving = (np.arange(n_boxes) % 2) == 0
        j, i = np.meshgrid(np.arange(cfg.image_width, dtype=np.float64), np.arange(cfg.image_height, dtype=np.float64))
        self.dir_cam = np.stack([(j - cfg.cx) / cfg.fx, (i - cfg.cy) / cfg.fy, np.ones_like(j)], -1)

    def pose(self, t: int):
        pos = np.array([0.0, 0.0, self.cfg.speed * t], np.float64)
        yaw = 0.02 * np.sin(0.1 * t)
        quat = np.array([np.cos(yaw / 2), 0.0, np.sin(yaw / 2), 0.0], np.float64)
        return pos, quat, yaw

    def box_min(self, k: int, t: int):
        x = self.box_x0[k] + (BOX_SPEED_X * t if self.box_moving[k] else 0.0)
        return np.array([x, 1.5 - self.box_size[1], self.box_z0[k]])

    def frame(self, t: int) -> Frame:
        cfg = self.cfg
        pos, quat, yaw = self.pose(t)
        c, s = np.cos(yaw), np.sin(yaw)
        R = np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])
        d = self.dir_cam @ R.T                                           # world ray per unit camera depth
        H, W = cfg.image_height, cfg.image_width
        depth = np.full((H, W), np.inf)
        cls = np.full((H, W), V_TERRAIN, np.uint8)
        owner = np.full((H, W), -1, np.int32)

        def hit(sd, value, own=-1):
            nonlocal depth, cls, owner
            m = (sd > 0) & (sd < depth)
            depth = np.where(m, sd, depth)
            cls = np.where(m, np.uint8(value), cls)
            owner = np.where(m, own, owner)

        with np.errstate(divide="ignore", invalid="ignore"):
            hit(np.where(d[..., 1] > 1e-9, (1.5 - pos[1]) / d[..., 1], np.inf), V_ROAD)
            hit(np.where(d[..., 0] > 1e-9, (self.wall_x - pos[0]) / d[..., 0], np.inf), V_BUILDING)
            hit(np.where(d[..., 0] < -1e-9, (-self.wall_x - pos[0]) / d[..., 0], np.inf), V_BUILDING)
            hit(np.full((H, W), self.back_depth), V_WALL)
            for k in range(self.n_boxes):
                lo = self.box_min(k, t)
                hi = lo + self.box_size
                if lo[2] - pos[2] > 1.5 * cfg.depth_max or hi[2] < pos[2]:
                    continue
                t0 = (lo - pos) / d
                t1 = (hi - pos) / d
                tn = np.minimum(t0, t1).max(-1)
                tf = np.maximum(t0, t1).min(-1)
                hit(np.where((tn <= tf) & (tn > 0), tn, np.inf), V_VEHICLE, k)

        rng = np.random.default_rng([cfg.seed, t])
        if self.noise:
            depth = depth * rng.normal(1.0, 0.01, (H, W))
            u = rng.random((H, W))
            depth = np.where(u < 0.02, cfg.depth_max + 1.0, depth)       # "far" pixels
            depth = np.where((u >= 0.02) & (u < 0.03), 0.0, depth)        # invalid-low pixels
        depth = np.where(np.isfinite(depth), depth, cfg.depth_max + 1.0).astype(np.float32)

        masks = [MaskKpts(track_id=65535, label="static", mask=cls)]
        moves = []
        if self.instances:
            for k in range(self.n_boxes):
                mk = owner == k
                if not mk.any():
                    continue
                lo = self.box_min(k, t)
                corners = np.array([[lo[0] + a * self.box_size[0], lo[1] + b * self.box_size[1], lo[2] + e * self.box_size[2]]
                                    for a in (0, 1) for b in (0, 1) for e in (0, 1)])
                masks.append(MaskKpts(track_id=k + 1, label="Vehicle", mask=(mk * np.uint8(255)).astype(np.uint8), kpts_current=corners))
                if self.box_moving[k] and t > 0:
                    T = np.eye(4, dtype=np.float32)
                    T[0, 3] = BOX_SPEED_X
                    moves.append((k + 1, T))
        return Frame(t, depth, masks, pos, quat, moves, [])
