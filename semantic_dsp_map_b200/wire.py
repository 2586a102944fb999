"""ROS-free ingest of the reference node's wire format.

The reference's node (src/mapping.cpp:142-156) synchronises three ROS 1 topics and turns them into one ``update()`` call:
``sensor_msgs/Image`` (depth, 32FC1), ``geometry_msgs/PoseStamped`` (camera pose) and ``mask_kpts_msgs/MaskGroup``
(``std_msgs/Header header; MaskKpts[] objects`` with ``MaskKpts = string label, uint16 track_id, sensor_msgs/Image mask,
Keypoint bbox_tl, bbox_br, Keypoint[] kpts_curr, kpts_last`` and ``Keypoint = float32 x, y, z``;
script/mmseg/mask_kpts_msgs/msg/_MaskKpts.py:17-24, _MaskGroup.py:17-18).  This module decodes (and, for tests and recorded
replays, encodes) the ROS 1 serialisation of those messages - little-endian scalars, ``uint32`` length-prefixed strings and
arrays - and applies the node's conversion rules (``syncCallback``, src/mapping.cpp:165-260) so that recorded traffic can be
replayed through the C ABI without ROS, cv_bridge or the generated message classes:

* frames whose depth ``header.seq`` is < 2 are dropped (:172); depth and mask group must carry the same ``seq`` (:194-200);
* masks are taken as MONO8 (:212), depth as 32FC1 (:180); the bounding box and the key points are read only for non-static
  objects (:214-241); an object whose current / last key-point counts differ is skipped (:244-247).

``tests/golden/wire_maskgroup.bin`` was produced by the reference's own generated serialiser
(tests/golden/make_wire_golden.py); tests/test_wire_cpu.py decodes it and re-encodes it byte for byte.
"""
from __future__ import annotations

import struct
from dataclasses import dataclass, field
from typing import List, Optional, Tuple

import numpy as np

from .map import MaskKpts

_ENCODINGS = {"32FC1": (np.float32, 1), "mono8": (np.uint8, 1), "8UC1": (np.uint8, 1), "16UC1": (np.uint16, 1), "mono16": (np.uint16, 1)}


class WireError(ValueError):
    pass


class _Reader:
    def __init__(self, buf: bytes, pos: int = 0):
        self.buf, self.pos = memoryview(buf), pos

    def take(self, fmt: str):
        st = struct.Struct("<" + fmt)
        if self.pos + st.size > len(self.buf):
            raise WireError("buffer underfill")
        out = st.unpack_from(self.buf, self.pos)
        self.pos += st.size
        return out

    def blob(self) -> bytes:
        (n,) = self.take("I")
        if self.pos + n > len(self.buf):
            raise WireError("buffer underfill")
        out = bytes(self.buf[self.pos:self.pos + n])
        self.pos += n
        return out

    def string(self) -> str:
        return self.blob().decode("utf-8")


@dataclass
class Header:
    seq: int = 0
    secs: int = 0
    nsecs: int = 0
    frame_id: str = ""

    @property
    def stamp(self) -> float:
        return self.secs + 1e-9 * self.nsecs


@dataclass
class Image:
    header: Header
    height: int
    width: int
    encoding: str
    is_bigendian: int
    step: int
    data: bytes

    def array(self, want: str) -> np.ndarray:
        """What cv_bridge::toCvCopy(msg, want) yields for the encodings the node uses (no colour conversion is needed:
        depth arrives as 32FC1, masks as mono8 / 8UC1)."""
        if self.encoding not in _ENCODINGS:
            raise WireError(f"unsupported image encoding {self.encoding!r}")
        dtype, _ = _ENCODINGS[self.encoding]
        wdtype, _ = _ENCODINGS[want]
        if np.dtype(dtype) != np.dtype(wdtype):
            raise WireError(f"image is {self.encoding}, the node reads it as {want}")
        if self.step < self.width * np.dtype(dtype).itemsize or len(self.data) < self.step * self.height:
            raise WireError("image data shorter than step * height")
        dt = np.dtype(dtype).newbyteorder(">" if self.is_bigendian else "<")
        rows = np.frombuffer(self.data, np.uint8, self.step * self.height).reshape(self.height, self.step)
        return np.ascontiguousarray(rows[:, : self.width * dt.itemsize]).view(dt).astype(dtype, copy=False).reshape(self.height, self.width)


def _read_header(r: _Reader) -> Header:
    seq, secs, nsecs = r.take("3I")
    return Header(seq, secs, nsecs, r.string())


def _read_image(r: _Reader) -> Image:
    h = _read_header(r)
    height, width = r.take("2I")
    enc = r.string()
    big, step = r.take("BI")
    return Image(h, height, width, enc, big, step, r.blob())


def decode_image(buf: bytes) -> Image:
    """sensor_msgs/Image."""
    return _read_image(_Reader(buf))


def decode_pose_stamped(buf: bytes) -> Tuple[Header, np.ndarray, np.ndarray]:
    """geometry_msgs/PoseStamped -> (header, position xyz, orientation as (w, x, y, z) like Eigen::Quaterniond's constructor,
    src/mapping.cpp:185-186)."""
    r = _Reader(buf)
    h = _read_header(r)
    px, py, pz, qx, qy, qz, qw = r.take("7d")
    return h, np.array([px, py, pz]), np.array([qw, qx, qy, qz])


@dataclass
class WireObject:
    label: str
    track_id: int
    mask: Image
    bbox_tl: Tuple[float, float, float]
    bbox_br: Tuple[float, float, float]
    kpts_curr: np.ndarray   # (n, 3) float32
    kpts_last: np.ndarray


@dataclass
class MaskGroup:
    header: Header
    objects: List[WireObject] = field(default_factory=list)


def decode_mask_group(buf: bytes) -> MaskGroup:
    r = _Reader(buf)
    g = MaskGroup(_read_header(r))
    (n,) = r.take("I")
    for _ in range(n):
        label = r.string()
        (track,) = r.take("H")
        mask = _read_image(r)
        tl, br = r.take("3f"), r.take("3f")
        kp = []
        for _k in range(2):
            (m,) = r.take("I")
            kp.append(np.asarray(r.take(f"{3 * m}f"), np.float32).reshape(m, 3))
        g.objects.append(WireObject(label, track, mask, tl, br, kp[0], kp[1]))
    if r.pos != len(r.buf):
        raise WireError("trailing bytes after MaskGroup")
    return g


# ---- encoders (tests, recorded replays) ---------------------------------------------------------------------------------
def _w_string(s: str) -> bytes:
    b = s.encode("utf-8")
    return struct.pack("<I", len(b)) + b


def _w_header(h: Header) -> bytes:
    return struct.pack("<3I", h.seq, h.secs, h.nsecs) + _w_string(h.frame_id)


def encode_image(img: Image) -> bytes:
    return (_w_header(img.header) + struct.pack("<2I", img.height, img.width) + _w_string(img.encoding) + struct.pack("<BI", img.is_bigendian, img.step)
            + struct.pack("<I", len(img.data)) + bytes(img.data))


def image_from_array(a: np.ndarray, encoding: str, header: Optional[Header] = None) -> Image:
    a = np.ascontiguousarray(a)
    return Image(header or Header(), a.shape[0], a.shape[1], encoding, 0, a.shape[1] * a.dtype.itemsize, a.tobytes())


def encode_pose_stamped(h: Header, position, orientation_wxyz) -> bytes:
    w, x, y, z = [float(v) for v in orientation_wxyz]
    return _w_header(h) + struct.pack("<7d", *[float(v) for v in position], x, y, z, w)


def encode_mask_group(g: MaskGroup) -> bytes:
    out = [_w_header(g.header), struct.pack("<I", len(g.objects))]
    for o in g.objects:
        out += [_w_string(o.label), struct.pack("<H", o.track_id), encode_image(o.mask), struct.pack("<3f", *o.bbox_tl), struct.pack("<3f", *o.bbox_br)]
        for kp in (o.kpts_curr, o.kpts_last):
            kp = np.asarray(kp, np.float32).reshape(-1, 3)
            out += [struct.pack("<I", len(kp)), kp.tobytes()]
    return b"".join(out)


# ---- the node's conversion (syncCallback) ---------------------------------------------------------------------------------
@dataclass
class SyncedFrame:
    depth: np.ndarray
    tracking_result: List[MaskKpts]
    position: np.ndarray
    orientation: np.ndarray
    time_stamp: float
    seq: int


def to_frame(depth_msg: bytes, pose_msg: bytes, mask_group_msg: bytes) -> Optional[SyncedFrame]:
    """src/mapping.cpp:165-260.  Returns None where the node returns without calling update()."""
    depth_img = decode_image(depth_msg)
    pose_h, position, orientation = decode_pose_stamped(pose_msg)
    if depth_img.header.seq < 2:                                   # :172
        return None
    depth = depth_img.array("32FC1")                               # :180
    group = decode_mask_group(mask_group_msg)
    if depth_img.header.seq != group.header.seq:                   # :194-200
        return None
    result = []
    for o in group.objects:                                        # :205-251 (a uint16 track id is always in [0, 65535])
        mk = MaskKpts(track_id=o.track_id, label=o.label, mask=o.mask.array("mono8"))
        if o.label != "static":
            mk.bbox = (int(o.bbox_tl[0]), int(o.bbox_tl[1]), int(o.bbox_br[0]), int(o.bbox_br[1]))   # float -> int fields of BBox2D
            mk.kpts_current = o.kpts_curr.astype(np.float64)
            mk.kpts_previous = o.kpts_last.astype(np.float64)
            if len(mk.kpts_current) != len(mk.kpts_previous):      # :244-247
                continue
        result.append(mk)
    return SyncedFrame(depth, result, position, orientation, pose_h.stamp, depth_img.header.seq)


def write_record(f, depth_msg: bytes, pose_msg: bytes, mask_group_msg: bytes):
    """One synchronised triple of a recorded replay file: three uint32-length-prefixed serialised messages."""
    for b in (depth_msg, pose_msg, mask_group_msg):
        f.write(struct.pack("<I", len(b)))
        f.write(b)


def read_records(f):
    while True:
        head = f.read(4)
        if len(head) < 4:
            return
        out = [f.read(struct.unpack("<I", head)[0])]
        for _ in range(2):
            (n,) = struct.unpack("<I", f.read(4))
            out.append(f.read(n))
        yield tuple(out)


def replay(dsp_map, records, if_get_freespace=False, moves_for=None):
    """Feed recorded (depth, pose, mask group) triples through ``SemanticDSPMap.update`` the way the node does; yields
    (seq, OccupancyResult).  ``moves_for(frame) -> (moves, remove_tracks)`` replaces the built-in host object layer with the
    caller's own; by default the map runs objectLevelUpdate itself from the key points and the pose time stamp, like the node."""
    for rec in records:
        fr = to_frame(*rec)
        if fr is None:
            continue
        moves, remove = moves_for(fr) if moves_for else (None, ())
        yield fr.seq, dsp_map.update(fr.depth, fr.tracking_result, fr.position, fr.orientation, if_get_freespace, fr.time_stamp, moves=moves,
                                     remove_tracks=remove)
