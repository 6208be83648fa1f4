"""The C-ABI library loads on a CPU-only box and exports every symbol include/cvgraft.h declares; without a
GPU it fails loudly instead of falling back."""
import os
import re

import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def test_library_exports_every_declared_symbol():
    from computervision_objectdetection_featurematching_b200 import _lib, build
    build.build()
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "cvgraft.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(cvg_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in cvgraft.h but not exported"
        assert name in _lib.SYMBOLS, f"{name} has no ctypes prototype in _lib.py"
    assert lib.cvg_version().startswith(b"cvgraft")


def test_param_defaults_match_reference_constants():
    from computervision_objectdetection_featurematching_b200 import api
    p = api.detect_params()                     # src/TestsDetector.cpp:21-25
    assert abs(p.ratio - 0.9) < 1e-7 and p.min_inliers == 4
    assert abs(p.det_lo - 0.1) < 1e-7 and p.det_hi == 10.0
    assert p.ransac.threshold == 5.0 and p.ransac.max_iters == 2000 and p.ransac.confidence == 0.995


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from computervision_objectdetection_featurematching_b200 import api
    with pytest.raises(api.CvgError) as e:
        api.Context(0)
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)


def test_no_cpu_fallback_for_multi_device_contexts():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from computervision_objectdetection_featurematching_b200 import api
    with pytest.raises(api.CvgError) as e:
        api.Context([0, 1])                                     # cvg_create_multi
    assert e.value.code == 2 and "no CPU fallback" in str(e.value)
    lib = api._lib.load()
    assert lib.cvg_num_devices(None) == 0 and lib.cvg_exchange_kind(None) == b"none"
    assert lib.cvg_set_lanes(None, 2) == 1                     # CVG_ERR_INVALID, no crash on a NULL context


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "computervision_objectdetection_featurematching_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "cvoracle" not in txt and "oracle/" not in txt, f
