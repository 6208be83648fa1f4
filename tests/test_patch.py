"""patches/reference_cvgraft.patch applies to the reference as it lies under /root/reference (build container only) and puts
the three hooks where INTEGRATION.md says: the CMake option, the upload after the model loop, the fused call at the head of
the view loop."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
PATCH = os.path.join(ROOT, "patches", "reference_cvgraft.patch")
REF = "/root/reference"


def test_patch_touches_only_the_three_seam_files():
    files = [l.split()[1][2:] for l in open(PATCH) if l.startswith("+++ b/")]
    assert sorted(files) == ["CMakeLists.txt", "src/ModelsDetector.cpp", "src/TestsDetector.cpp"]
    added = [l[1:] for l in open(PATCH) if l.startswith("+") and not l.startswith("+++")]
    removed = [l for l in open(PATCH) if l.startswith("-") and not l.startswith("---")]
    assert not removed                                           # nothing of the reference is deleted: the option OFF build is unchanged
    assert any("option(USE_CVGRAFT" in l for l in added)
    assert any("cvg::uploadModels(models, cvg::residentModels())" in l for l in added)
    assert any("cvg::detectAtScale(cvg::residentModels()" in l for l in added)


def test_patch_applies_to_the_reference(tmp_path):
    if not os.path.isdir(os.path.join(REF, "src")):
        pytest.skip("reference sources not present (build container only)")
    work = tmp_path / "ref"
    os.makedirs(work / "src")
    shutil.copy(os.path.join(REF, "CMakeLists.txt"), work)
    for f in ("ModelsDetector.cpp", "TestsDetector.cpp"):
        shutil.copy(os.path.join(REF, "src", f), work / "src")
    subprocess.check_call(["git", "apply", "--check", PATCH], cwd=work)
    subprocess.check_call(["git", "apply", PATCH], cwd=work)
    td = open(work / "src" / "TestsDetector.cpp").read()
    head = td.index("cvg::detectAtScale(cvg::residentModels()")
    assert td.index("auto detectAtScale = [&]") < head < td.index("matcher.knnMatch(model.descriptors[i], sceneDesc, knnMatches, 2)")
    md = open(work / "src" / "ModelsDetector.cpp").read()
    assert md.index("models.push_back(model);") < md.index("cvg::uploadModels(models, cvg::residentModels())")
    # with the option off the preprocessor removes every added line of the two sources
    for name, text in (("TestsDetector.cpp", td), ("ModelsDetector.cpp", md)):
        stripped, skip = [], 0
        for line in text.splitlines():
            if line.strip() == "#ifdef USE_CVGRAFT":
                skip += 1
            elif line.strip() == "#endif" and skip:
                skip -= 1
            elif not skip:
                stripped.append(line)
        assert "\n".join(stripped).strip() == open(os.path.join(REF, "src", name)).read().strip(), name
