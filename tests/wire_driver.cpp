// Decodes a recorded (depth, pose, mask group) triple with adapter/mask_group_wire.h and prints what the node would hand to
// update(): used by tests/test_wire_cpu.py against tests/golden/wire_maskgroup.bin (bytes from the reference's own serialiser).
#include "../adapter/mask_group_wire.h"

#include <cstdio>
#include <fstream>
#include <iterator>

static std::vector<uint8_t> slurp(const char* path) {
    std::ifstream f(path, std::ios::binary);
    return std::vector<uint8_t>(std::istreambuf_iterator<char>(f), std::istreambuf_iterator<char>());
}

int main(int argc, char** argv) {
    if (argc < 4) return 2;
    dspwire::Frame fr;
    try {
        if (!dspwire::toFrame(slurp(argv[1]), slurp(argv[2]), slurp(argv[3]), fr)) { std::printf("dropped\n"); return 0; }
    } catch (const std::exception& e) { std::printf("error %s\n", e.what()); return 0; }
    std::printf("seq %u stamp %.3f depth %dx%d d00 %.3f pos %.1f %.1f %.1f quat %.1f %.1f %.1f %.1f\n", fr.seq, fr.time_stamp, fr.depth.rows, fr.depth.cols,
                fr.depth.at<float>(0, 0), fr.position.x(), fr.position.y(), fr.position.z(), fr.orientation.w(), fr.orientation.x(), fr.orientation.y(),
                fr.orientation.z());
    for (const MaskKpts& m : fr.tracking_result) {
        unsigned sum = 0;
        for (int i = 0; i < m.mask.rows; ++i) for (int j = 0; j < m.mask.cols; ++j) sum += m.mask.at<unsigned char>(i, j);
        std::printf("object %s %d mask %dx%d sum %u kpts %zu\n", m.label.c_str(), m.track_id, m.mask.rows, m.mask.cols, sum, m.kpts_current.size());
    }
    return 0;
}
