"""ROS-free ingest of the node's wire format (semantic_dsp_map_b200/wire.py).

tests/golden/wire_maskgroup.bin is a mask_kpts_msgs/MaskGroup serialised by the reference's own generated message code
(tests/golden/make_wire_golden.py): the decoder must read it, the encoder must reproduce it byte for byte, and the
conversion must follow the node's rules (src/mapping.cpp:165-260)."""
import io
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import make_wire_golden as G  # noqa: E402

from semantic_dsp_map_b200 import wire  # noqa: E402

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "wire_maskgroup.bin")


def test_decodes_the_reference_serialisation_and_reencodes_it_byte_for_byte():
    buf = open(GOLDEN, "rb").read()
    ex = G.example()
    g = wire.decode_mask_group(buf)
    assert (g.header.seq, g.header.secs, g.header.nsecs, g.header.frame_id) == (ex["seq"], ex["secs"], ex["nsecs"], ex["frame_id"])
    assert len(g.objects) == len(ex["objects"])
    for o, e in zip(g.objects, ex["objects"]):
        assert (o.label, o.track_id) == (e["label"], e["track_id"])
        np.testing.assert_array_equal(o.mask.array("mono8"), e["mask"])
        assert tuple(o.bbox_tl) == tuple(map(float, e["tl"])) and tuple(o.bbox_br) == tuple(map(float, e["br"]))
        np.testing.assert_array_equal(o.kpts_curr, e["kc"])
        np.testing.assert_array_equal(o.kpts_last, e["kl"])
    assert wire.encode_mask_group(g) == buf
    with pytest.raises(wire.WireError):
        wire.decode_mask_group(buf[:-3])
    with pytest.raises(wire.WireError):
        wire.decode_mask_group(buf + b"\0")


def test_conversion_follows_the_nodes_rules():
    buf = open(GOLDEN, "rb").read()
    g = wire.decode_mask_group(buf)
    depth = np.linspace(0.5, 9.0, 12 * 20, dtype=np.float32).reshape(12, 20)

    def triple(depth_seq, group_seq=None):
        h = wire.Header(depth_seq, 1700000000, 5, "zed2")
        gg = wire.MaskGroup(wire.Header(depth_seq if group_seq is None else group_seq, 1, 2, "zed2"), g.objects)
        return (wire.encode_image(wire.image_from_array(depth, "32FC1", h)),
                wire.encode_pose_stamped(wire.Header(depth_seq, 1700000001, 500000000, "map"), [1.0, 2.0, 3.0], [0.5, 0.1, 0.2, 0.3]), wire.encode_mask_group(gg))

    assert wire.to_frame(*triple(1)) is None                     # the first two depth images are ignored (:172)
    assert wire.to_frame(*triple(5, 6)) is None                  # sequence mismatch (:194-200)
    fr = wire.to_frame(*triple(7))
    np.testing.assert_array_equal(fr.depth, depth)
    np.testing.assert_array_equal(fr.position, [1.0, 2.0, 3.0])
    np.testing.assert_array_equal(fr.orientation, [0.5, 0.1, 0.2, 0.3])     # (w, x, y, z), :186
    assert abs(fr.time_stamp - 1700000001.5) < 1e-6
    # the static mask and the vehicle survive; "person" has 1 current and 0 last key points -> skipped (:244-247)
    assert [(m.label, m.track_id) for m in fr.tracking_result] == [("static", 65535), ("Vehicle", 3)]
    assert len(fr.tracking_result[0].kpts_current) == 0           # key points are not read for the static mask (:214)
    assert fr.tracking_result[1].bbox == (5, 3, 13, 8) and fr.tracking_result[1].kpts_current.shape == (4, 3)


def test_record_file_round_trip_and_image_errors():
    depth = np.ones((4, 6), np.float32)
    img = wire.encode_image(wire.image_from_array(depth, "32FC1", wire.Header(9)))
    pose = wire.encode_pose_stamped(wire.Header(9), [0, 0, 0], [1, 0, 0, 0])
    grp = wire.encode_mask_group(wire.MaskGroup(wire.Header(9)))
    f = io.BytesIO()
    for _ in range(3):
        wire.write_record(f, img, pose, grp)
    f.seek(0)
    recs = list(wire.read_records(f))
    assert len(recs) == 3 and recs[1] == (img, pose, grp)
    with pytest.raises(wire.WireError):
        wire.decode_image(img).array("mono8")                    # a 32FC1 image is not a mask
    bad = wire.image_from_array(depth, "32FC1")
    bad.encoding = "bgr8"
    with pytest.raises(wire.WireError):
        bad.array("32FC1")


def test_cpp_decoder_of_the_adapter_reads_the_same_bytes(tmp_path):
    """adapter/mask_group_wire.h (the C++ twin an integrator would use) on the reference-serialised message."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "wire_driver")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-w", "-I" + os.path.join(root, "oracle", "ref_stubs"), "-I" + os.path.join(root, "include"),
                           "-I" + os.path.join(root, "adapter"), os.path.join(root, "tests", "wire_driver.cpp"), "-o", exe,
                           "-L" + os.path.join(root, "semantic_dsp_map_b200"), "-l:libdspmap_b200.so", "-Wl,-rpath," + os.path.join(root, "semantic_dsp_map_b200")])
    group = open(GOLDEN, "rb").read()
    depth = np.full((12, 20), 2.5, np.float32)

    def run(depth_seq):
        files = {}
        for name, blob in (("depth", wire.encode_image(wire.image_from_array(depth, "32FC1", wire.Header(depth_seq, 10, 0, "zed2")))),
                           ("pose", wire.encode_pose_stamped(wire.Header(depth_seq, 1700000001, 500000000, "map"), [1.0, 2.0, 3.0], [0.5, 0.1, 0.2, 0.3])), ("group", group)):
            files[name] = str(tmp_path / name)
            open(files[name], "wb").write(blob)
        return subprocess.run([exe, files["depth"], files["pose"], files["group"]], capture_output=True, text=True, check=True).stdout.strip().splitlines()

    assert run(1) == ["dropped"] and run(9) == ["dropped"]                     # seq < 2; seq mismatch (the golden group has seq 7)
    out = run(7)
    ex = G.example()
    assert out[0] == "seq 7 stamp 1700000001.500 depth 12x20 d00 2.500 pos 1.0 2.0 3.0 quat 0.5 0.1 0.2 0.3"
    assert out[1] == "object static 65535 mask 12x20 sum %d kpts 0" % int(ex["objects"][0]["mask"].sum())
    assert out[2] == "object Vehicle 3 mask 12x20 sum %d kpts 4" % int(ex["objects"][1]["mask"].sum())
    assert len(out) == 3                                                        # "person": key-point counts differ -> skipped
