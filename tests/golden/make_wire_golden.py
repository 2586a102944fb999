"""tests/golden/wire_maskgroup.bin: a mask_kpts_msgs/MaskGroup serialised by the REFERENCE'S OWN generated message code
(script/mmseg/mask_kpts_msgs/msg/_MaskGroup.py, imported from /root/reference; genpy / std_msgs / sensor_msgs, absent here,
are stood in by a few plain classes with the attributes that code reads).  Build container only:

    python tests/golden/make_wire_golden.py
"""
import io
import os
import struct
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_MSGS = "/root/reference/script/mmseg"


def install_stand_ins():
    genpy = types.ModuleType("genpy")

    class Message:
        __slots__ = []

        def __init__(self, *args, **kwds):
            for k in self.__slots__:
                setattr(self, k, None)
            for k, v in zip(self.__slots__, args):
                setattr(self, k, v)
            for k, v in kwds.items():
                setattr(self, k, v)

    class Time:
        def __init__(self, secs=0, nsecs=0):
            self.secs, self.nsecs = secs, nsecs

    genpy.Message, genpy.Time, genpy.struct_I = Message, Time, struct.Struct("<I")
    genpy.DeserializationError = type("DeserializationError", (Exception,), {})
    sys.modules["genpy"] = genpy
    std = types.ModuleType("std_msgs"); std.msg = types.ModuleType("std_msgs.msg")

    class Header:
        def __init__(self, seq=0, secs=0, nsecs=0, frame_id=""):
            self.seq, self.stamp, self.frame_id = seq, Time(secs, nsecs), frame_id

    std.msg.Header = Header
    sys.modules["std_msgs"], sys.modules["std_msgs.msg"] = std, std.msg
    sen = types.ModuleType("sensor_msgs"); sen.msg = types.ModuleType("sensor_msgs.msg")

    class Image:
        def __init__(self, header=None, height=0, width=0, encoding="", is_bigendian=0, step=0, data=b""):
            self.header, self.height, self.width, self.encoding = header or Header(), height, width, encoding
            self.is_bigendian, self.step, self.data = is_bigendian, step, data

    sen.msg.Image = Image
    sys.modules["sensor_msgs"], sys.modules["sensor_msgs.msg"] = sen, sen.msg
    return Header, Image


def example():
    """The content of the golden message, as plain Python (also what tests/test_wire_cpu.py expects to decode)."""
    rng = np.random.default_rng(5)
    static = rng.integers(0, 12, (12, 20), dtype=np.uint8)
    car = np.zeros((12, 20), np.uint8); car[3:8, 5:13] = 255
    ped = np.zeros((12, 20), np.uint8); ped[6:11, 15:18] = 255
    return dict(seq=7, secs=1700000000, nsecs=123456789, frame_id="zed2",
                objects=[dict(label="static", track_id=65535, mask=static, tl=(0, 0, 0), br=(0, 0, 0), kc=np.zeros((0, 3), np.float32), kl=np.zeros((0, 3), np.float32)),
                         dict(label="Vehicle", track_id=3, mask=car, tl=(5, 3, 0), br=(13, 8, 0),
                              kc=np.array([[1, 2, 9], [2.5, 2, 9], [1, 0.5, 9], [1, 2, 13]], np.float32), kl=np.array([[0.9, 2, 9], [2.4, 2, 9], [0.9, 0.5, 9], [0.9, 2, 13]], np.float32)),
                         dict(label="person", track_id=4, mask=ped, tl=(15, 6, 0), br=(18, 11, 0), kc=np.array([[4, 1, 6]], np.float32), kl=np.zeros((0, 3), np.float32))])


def main():
    Header, Image = install_stand_ins()
    sys.path.insert(0, REF_MSGS)
    from mask_kpts_msgs.msg import Keypoint, MaskGroup, MaskKpts   # the reference's generated classes
    ex = example()
    g = MaskGroup(header=Header(ex["seq"], ex["secs"], ex["nsecs"], ex["frame_id"]), objects=[])
    for o in ex["objects"]:
        img = Image(Header(ex["seq"], ex["secs"], ex["nsecs"], ex["frame_id"]), o["mask"].shape[0], o["mask"].shape[1], "mono8", 0, o["mask"].shape[1], o["mask"].tobytes())
        kp = lambda a: [Keypoint(x=float(p[0]), y=float(p[1]), z=float(p[2])) for p in a]  # noqa: E731
        g.objects.append(MaskKpts(label=o["label"], track_id=o["track_id"], mask=img, bbox_tl=Keypoint(x=o["tl"][0], y=o["tl"][1], z=o["tl"][2]),
                                  bbox_br=Keypoint(x=o["br"][0], y=o["br"][1], z=o["br"][2]), kpts_curr=kp(o["kc"]), kpts_last=kp(o["kl"])))
    buf = io.BytesIO()
    g.serialize(buf)
    path = os.path.join(HERE, "wire_maskgroup.bin")
    open(path, "wb").write(buf.getvalue())
    print(path, len(buf.getvalue()), "bytes, md5sum of the message type", MaskGroup._md5sum)


if __name__ == "__main__":
    main()
