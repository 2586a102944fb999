"""cv::cvtColor(COLOR_RGB2HSV / COLOR_HSV2RGB) on 8-bit pixels, as getOccupancyResult uses it (semantic_dsp_map.h:1335-1350):
the oracle's restatement and the stand-in inside oracle/ref_stubs (the one the compiled reference runs on) against the
real OpenCV of this image (cv2), on ALL 2^24 inputs each way."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle_lib import load_oracle
from ref_lib import have_ref, ref_lib_path

cv2 = pytest.importorskip("cv2")


def all_triples():
    a = np.arange(256, dtype=np.uint8)
    x, y, z = np.meshgrid(a, a, a, indexing="ij")
    return np.ascontiguousarray(np.stack([x, y, z], -1).reshape(-1, 3))


def test_oracle_colour_conversions_match_opencv_exhaustively():
    lib = load_oracle()
    src = all_triples()
    out = np.empty_like(src)
    lib.orc_rgb_to_hsv(src.ctypes.data, out.ctypes.data, len(src))
    np.testing.assert_array_equal(out, cv2.cvtColor(src.reshape(1, -1, 3), cv2.COLOR_RGB2HSV)[0])
    lib.orc_hsv_to_rgb(src.ctypes.data, out.ctypes.data, len(src))
    want = cv2.cvtColor(src.reshape(1, -1, 3), cv2.COLOR_HSV2RGB)[0]
    ok = src[:, 0] < 180                                    # H of an 8-bit HSV image is in [0, 180)
    np.testing.assert_array_equal(out[ok], want[ok])


@pytest.mark.skipif(not have_ref(3), reason="oracle/_ref not built")
def test_reference_build_stand_in_cvtcolor_matches_opencv_exhaustively():
    lib = C.CDLL(ref_lib_path(3))
    lib.ref_cvt_color.argtypes = [C.c_int32, C.c_void_p, C.c_void_p, C.c_uint32]
    src = all_triples()
    out = np.empty_like(src)
    lib.ref_cvt_color(0, src.ctypes.data, out.ctypes.data, len(src))
    np.testing.assert_array_equal(out, cv2.cvtColor(src.reshape(1, -1, 3), cv2.COLOR_RGB2HSV)[0])
    lib.ref_cvt_color(1, src.ctypes.data, out.ctypes.data, len(src))
    want = cv2.cvtColor(src.reshape(1, -1, 3), cv2.COLOR_HSV2RGB)[0]
    ok = src[:, 0] < 180
    np.testing.assert_array_equal(out[ok], want[ok])
