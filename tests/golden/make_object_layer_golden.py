"""tests/golden/object_layer_reference.txt: what the REFERENCE'S OWN object layer decides, frame by frame, on the scenarios of
tests/object_layer/scenarios.h (tests/object_layer/reference_harness.cpp: objectLevelUpdate / ObjectSet / Kabsch-RANSAC of the
unmodified headers driven through SemanticDSPMap::update; Eigen / OpenCV / PCL from oracle/ref_stubs, whose JacobiSVD is a
one-sided Jacobi SVD).  SETTING 3 compiles the tree as it is; SETTING 2 reaches the headers through a throw-away directory
of symlinks in which only settings/settings.h has its "#define SETTING" line changed (the trick of oracle/build_ref.sh).
Build container only:

    python tests/golden/make_object_layer_golden.py
"""
import os
import re
import subprocess
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
SRC = os.path.join(ROOT, "tests", "object_layer")
LABELS_CSV = ("label_id,label,instance_id(-1 means not fixed),color_b,color_g,color_r\n0,background,65535,0,0,0\n1,road,65534,128,64,128\n"
              "12,person,-1,60,20,220\n14,Vehicle,-1,142,0,0\n")


def scenarios():
    """(name, setting) of every scenario, read off scenarios.h"""
    text = open(os.path.join(SRC, "scenarios.h")).read()
    return [(m.group(1), int(m.group(3) or 3)) for m in re.finditer(r'Scenario s\{"(\w+)", \d+, \{\}(, (\d))?\}', text)]


def include_dir(td, setting):
    if setting == 3:
        return os.path.join(REF, "include")
    inc = os.path.join(td, f"inc{setting}")
    for dp, _, fs in os.walk(os.path.join(REF, "include")):
        rel = os.path.relpath(dp, os.path.join(REF, "include"))
        os.makedirs(os.path.join(inc, rel), exist_ok=True)
        for f in fs:
            if os.path.join(rel, f) != os.path.join("settings", "settings.h"):
                os.symlink(os.path.join(dp, f), os.path.join(inc, rel, f))
    text = open(os.path.join(REF, "include", "settings", "settings.h")).read()
    new = re.sub(r"^#define SETTING 3 ", f"#define SETTING {setting} ", text, count=1, flags=re.M)
    assert new != text
    with open(os.path.join(inc, "settings", "settings.h"), "w") as f:
        f.write(new)
    return inc


def main():
    out = []
    with tempfile.TemporaryDirectory() as td:
        csv = os.path.join(td, "labels.csv")
        with open(csv, "w") as f:
            f.write(LABELS_CSV)
        exes = {}
        for name, setting in scenarios():
            if setting not in exes:
                exes[setting] = os.path.join(td, f"ol_ref_s{setting}")
                subprocess.check_call(["/usr/bin/g++", "-std=c++17", "-O2", "-march=x86-64-v3", "-ffp-contract=off", "-w", "-I" + os.path.join(ROOT, "oracle", "ref_stubs"),
                                       "-I" + include_dir(td, setting), "-I" + SRC, os.path.join(SRC, "reference_harness.cpp"), "-o", exes[setting]])
            r = subprocess.run([exes[setting], name, csv], stdout=subprocess.DEVNULL, stderr=subprocess.PIPE, text=True, check=True)
            lines = [ln for ln in r.stderr.splitlines() if ln.startswith("OBJ ")]
            assert lines, name
            out.append(f"SCENARIO {name}")
            out += lines
            print(name, "setting", setting, len(lines), "lines")
    with open(os.path.join(HERE, "object_layer_reference.txt"), "w") as f:
        f.write("\n".join(out) + "\n")


if __name__ == "__main__":
    main()
