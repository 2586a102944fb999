#!/bin/sh
# Builds oracle/_ref/libdsp_ref_s<SETTING>.so: the reference's own headers, compiled where they lie under
# $REF/include, against the stand-in third-party headers of oracle/ref_stubs (see ref_harness.cpp).
# TEST INFRASTRUCTURE.  Outputs go only into oracle/_ref/ (git-ignored; travels to the GPU box like any built .so).
#
#   sh oracle/build_ref.sh [SETTING ...]      default: 3 (the configuration the reference ships, settings.h:22)
#
# The name "B" builds BASELINE.json's benchmark configuration (ZED2 preset with 256 x 64 x 256 voxels and 8 slots per voxel,
# SURVEY.md section 8 config B): SETTING 3 with the four grid-size constants of that block changed - what docs/parameter_table.md
# tells a user to edit.  bench.py times it as the CPU baseline (cpu_baseline.kind = "reference").  "BV" is "B" with VERBOSE_MODE 1.
#
# SETTING 3 compiles the tree exactly as it is.  The reference selects its grid / camera preset with a hard
# "#define SETTING 3" (include/settings/settings.h:22) that no -D can override, so for another preset the headers are
# reached through a throw-away directory of SYMLINKS into $REF/include in which only settings/settings.h is a
# generated file with that one line changed; the directory is deleted after the compile and nothing is copied into
# the repository.
set -e
REF="${REF:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
# the system compiler: its libstdc++ is linked dynamically (the /opt/gcc toolchain of this image links it statically, and two
# static iostream copies in one process do not mix)
CXX="${REF_CXX:-/usr/bin/g++}"
# the reference's flags (CMakeLists.txt:5-8: -std=c++17 -O3 -ftree-vectorize -fPIC) with the oracle's two changes:
# -ffp-contract=off pins the floating-point order, -march=x86-64-v3 instead of -march=native because the .so travels
FLAGS="-std=c++17 -O3 -ftree-vectorize -march=x86-64-v3 -ffp-contract=off -fPIC -fno-access-control -shared -w"
[ -d "$REF/include" ] || { echo "build_ref: $REF/include not present, keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$HERE/_ref"
[ $# -gt 0 ] || set -- 3
for S in "$@"; do
    OUT="$HERE/_ref/libdsp_ref_s$S.so"
    if [ -f "$OUT" ] && [ "$OUT" -nt "$HERE/ref_harness.cpp" ] && [ "$OUT" -nt "$HERE/ref_stubs/Eigen/Dense" ] && \
       [ "$OUT" -nt "$HERE/ref_stubs/opencv2/opencv.hpp" ] && [ "$OUT" -nt "$HERE/ref_stubs/pcl/point_types.h" ]; then continue; fi
    if [ "$S" = "3" ]; then
        INC="$REF/include"
        TMP=""
    else
        TMP="$(mktemp -d)"
        (cd "$REF/include" && find . -type d) | while read -r d; do mkdir -p "$TMP/$d"; done
        (cd "$REF/include" && find . -type f) | while read -r f; do ln -s "$REF/include/$f" "$TMP/$f"; done
        rm "$TMP/settings/settings.h"
        if [ "$S" = "B" ] || [ "$S" = "BV" ]; then
            # BV: config B with the reference's own stage timers on (VERBOSE_MODE 1, settings.h:20; semantic_dsp_map.h:912-953
            # prints the six "Time used for ..." lines per frame): bench.py reads the CPU baseline's per-stage split from them
            VERB=0; [ "$S" != "BV" ] || VERB=1
            sed -e "s/^#define VERBOSE_MODE 0 /#define VERBOSE_MODE $VERB /" -e '/^#elif SETTING == 3/,/^#else/ s/C_VOXEL_NUM_AXIS_X_N = 7;/C_VOXEL_NUM_AXIS_X_N = 8;/' \
                -e '/^#elif SETTING == 3/,/^#else/ s/C_VOXEL_NUM_AXIS_Y_N = 5;/C_VOXEL_NUM_AXIS_Y_N = 6;/' \
                -e '/^#elif SETTING == 3/,/^#else/ s/C_VOXEL_NUM_AXIS_Z_N = 7;/C_VOXEL_NUM_AXIS_Z_N = 8;/' \
                -e '/^#elif SETTING == 3/,/^#else/ s/C_MAX_PARTICLE_NUM_PER_VOXEL_N = 2;/C_MAX_PARTICLE_NUM_PER_VOXEL_N = 3;/' \
                "$REF/include/settings/settings.h" > "$TMP/settings/settings.h"
            [ "$(grep -c 'C_VOXEL_NUM_AXIS_Y_N = 6;' "$TMP/settings/settings.h")" = "1" ]
            [ "$(grep -c 'C_MAX_PARTICLE_NUM_PER_VOXEL_N = 3;' "$TMP/settings/settings.h")" = "3" ]
            grep -q "^#define VERBOSE_MODE $VERB " "$TMP/settings/settings.h"
        else
            sed "s/^#define SETTING 3 /#define SETTING $S /" "$REF/include/settings/settings.h" > "$TMP/settings/settings.h"
            grep -q "^#define SETTING $S " "$TMP/settings/settings.h"
        fi
        INC="$TMP"
    fi
    echo "build_ref: SETTING $S -> $OUT"
    # SETTING 0 declares a 3.2 GB static particle array (docs/trouble_shooting.md:3-18): it only links with the medium code model
    MODEL=""; [ "$S" != "0" ] || MODEL="-mcmodel=medium"
    $CXX $FLAGS $MODEL -I"$HERE/ref_stubs" -I"$INC" -o "$OUT" "$HERE/ref_harness.cpp"
    [ -z "$TMP" ] || rm -rf "$TMP"
done
