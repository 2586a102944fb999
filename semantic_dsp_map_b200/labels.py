"""Label tables of the reference, evaluated on the host into the LUTs the C-ABI takes.

Mirrors the label globals of the reference (include/utils/data_base.h:108-212) and
``ObjectInfoHandler::readObjectInfo`` (include/utils/object_info_handler.h:28-91), including the
``std::unordered_map::operator[]`` behaviour the hot path depends on: a missing key is
default-inserted and reads as 0 / "" (SURVEY.md §7.3-e).
"""
from __future__ import annotations

import csv
from dataclasses import dataclass, field

import numpy as np

# include/utils/data_base.h:108-196 — the built-in tables used when no CSV is read
_DEFAULT_ROWS = [
    (0, "Background", 65535), (2, "Terrain", 65534), (3, "Sky", 65533), (4, "Tree", 65532),
    (5, "Vegetation", 65531), (6, "Building", 65530), (7, "Road", 65529), (8, "GuardRail", 65528),
    (9, "TrafficSign", 65527), (10, "TrafficLight", 65526), (11, "Pole", 65525), (12, "Misc", 65524),
    (13, "Truck", -1), (14, "Car", -1), (15, "Person", -1),
]
# include/utils/data_base.h:198-232 — g_label_color_map_default of the built-in tables, label id -> (b, g, r)
_DEFAULT_COLORS = {0: (0, 0, 0), 2: (200, 0, 210), 3: (255, 200, 90), 4: (0, 199, 0), 5: (0, 240, 90), 6: (140, 140, 140), 7: (100, 60, 100),
                   8: (255, 100, 250), 9: (0, 255, 255), 10: (0, 200, 200), 11: (0, 130, 255), 12: (80, 80, 80), 13: (60, 60, 160), 14: (80, 127, 255),
                   15: (139, 139, 0)}

# cfg/object_info_zed2.csv of the reference (label_id, label, instance_id); colours are not on the hot path
ZED2_ROWS = [
    (0, "background", 65535), (1, "road", 65534), (2, "sidewalk", 65533), (3, "building", 65532),
    (4, "wall", 65531), (5, "fence", 65530), (6, "pole", 65529), (7, "traffic_light", 65528),
    (8, "traffic_sign", 65527), (9, "vegetation", 65526), (10, "terrain", 65525), (11, "sky", 65524),
    (12, "person", -1), (13, "rider", 65522), (14, "Vehicle", -1), (15, "truck", 65521), (16, "bus", 65520),
    (17, "train", 65519), (18, "motorcycle", 65518), (19, "bicycle", 65517),
]


@dataclass
class LabelTables:
    """State of the reference's label globals after (optionally) reading a CSV."""
    label_id_map: dict = field(default_factory=dict)             # g_label_id_map_default
    label_id_map_reversed: dict = field(default_factory=dict)    # g_label_id_map_reversed
    label_id_map_static: dict = field(default_factory=dict)      # g_label_id_map_static
    label_to_instance: dict = field(default_factory=dict)        # g_label_to_instance_id_map_default
    instance_to_label: dict = field(default_factory=dict)        # g_instance_id_to_label_map_default
    movable_label_ids: set = field(default_factory=set)          # g_movable_object_label_ids_set
    label_colors: dict = field(default_factory=dict)             # g_label_color_map_default: label id -> (b, g, r)
    max_movable_instance_id: int = 65523                         # g_max_movable_object_instance_id

    @classmethod
    def builtin(cls) -> "LabelTables":
        t = cls()
        for label_id, name, inst in _DEFAULT_ROWS:
            t.label_id_map[name] = label_id
            t.label_id_map_reversed[label_id] = name
            if inst > 0:
                t.label_id_map_static[name] = label_id
                t.label_to_instance[name] = inst
                t.instance_to_label[inst] = name
            else:
                t.movable_label_ids.add(label_id)
        t.max_movable_instance_id = 65523
        t.label_colors = dict(_DEFAULT_COLORS)
        return t

    @classmethod
    def from_rows(cls, rows) -> "LabelTables":
        """object_info_handler.h:51-86: rebuild every table from (label_id, label, instance_id[, color_b, color_g, color_r]) rows."""
        t = cls()
        max_movable = 65535
        for row in rows:
            label_id, name, inst = row[:3]
            if len(row) >= 6:
                t.label_colors[label_id] = (int(row[3]) & 0xFF, int(row[4]) & 0xFF, int(row[5]) & 0xFF)   # cv::Vec3b(b, g, r), :56
            t.label_id_map[name] = label_id
            t.label_id_map_reversed[label_id] = name
            if inst > 0:
                t.label_id_map_static[name] = label_id
                t.label_to_instance[name] = inst
                t.instance_to_label[inst] = name
                if inst - 1 < max_movable:
                    max_movable = inst - 1
            else:
                t.movable_label_ids.add(label_id)
        t.max_movable_instance_id = max_movable
        return t

    @classmethod
    def read_object_info(cls, path: str) -> "LabelTables":
        """ObjectInfoHandler::readObjectInfo: a non-.csv path keeps the built-in tables."""
        if path.rsplit(".", 1)[-1] != "csv":
            return cls.builtin()
        rows = []
        with open(path, newline="") as f:
            rd = csv.reader(f)
            next(rd)  # header line is skipped (object_info_handler.h:138)
            for r in rd:
                if r:
                    rows.append((int(r[0]), r[1], int(r[2])) + tuple(int(v) for v in r[3:6]))
        return cls.from_rows(rows)

    # ---- what the hot path evaluates per pixel, folded into LUTs -------------------------------
    def label_id_of(self, name: str) -> int:
        """g_label_id_map_default[name] with default-insert semantics."""
        return self.label_id_map.get(name, 0)

    def static_value_to_track(self) -> np.ndarray:
        """pointcloud_tools.h:137-138: track = g_label_to_instance_id_map_default[g_label_id_map_reversed[v+1]]."""
        lut = np.zeros(256, np.uint16)
        for v in range(256):
            name = self.label_id_map_reversed.get(v + 1, "")
            lut[v] = self.label_to_instance.get(name, 0) & 0xFFFF
        return lut

    def static_track_labels(self):
        """pointcloud_tools.h:277-279: label = g_label_id_map_static[g_instance_id_to_label_map_default[track]]."""
        tracks = sorted(self.instance_to_label)
        labels = [self.label_id_map_static.get(self.instance_to_label[t], 0) & 0xFF for t in tracks]
        return np.asarray(tracks, np.uint16), np.asarray(labels, np.uint8)

    def sky_track(self) -> int:
        """pointcloud_tools.h:236: g_label_to_instance_id_map_default["Sky"] (case-sensitive, 0 when absent)."""
        return self.label_to_instance.get("Sky", 0)
