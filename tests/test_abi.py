"""CPU-side checks of the drop-in boundary: the shared library loads without a GPU, exports every
symbol include/okp.h declares, validates its arguments before touching the device, and the ctypes
mirror agrees with the header."""
import ctypes
import os
import re

import numpy as np
import pytest

from object_keypoints_b200 import _abi, _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_functions():
    text = open(os.path.join(ROOT, 'include', 'okp.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(okp_[a-z0-9_]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    if _lib.needs_build():
        _lib.build()
    return _lib.lib()


def test_every_declared_symbol_is_exported(lib):
    declared = header_functions()
    assert declared, "no declarations parsed from include/okp.h"
    assert sorted(_lib.EXPORTS) == declared
    for name in declared:
        assert hasattr(lib, name), f"libokp.so does not export {name}"


def header_prototypes():
    """name -> list of parameter declarations, parsed from include/okp.h."""
    text = open(os.path.join(ROOT, 'include', 'okp.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    out = {}
    for name, params in re.findall(r'\b(okp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;', text, flags=re.S):
        params = params.strip()
        out[name] = [] if params in ('', 'void') else [p.strip() for p in params.split(',')]
    return out


def test_ctypes_signatures_agree_with_the_header(lib):
    """Every binding in _lib.py takes as many arguments as the prototype declares, pointers where the header has
    pointers and floating-point types where it has double / float (catches a drifted binding without a GPU)."""
    prototypes = header_prototypes()
    assert sorted(prototypes) == sorted(_lib.EXPORTS)
    for name, params in prototypes.items():
        argtypes = getattr(lib, name).argtypes
        assert argtypes is not None, f"{name}: no argtypes declared"
        assert len(argtypes) == len(params), f"{name}: header has {len(params)} parameters, the binding {len(argtypes)}"
        for declaration, ctype in zip(params, argtypes):
            is_pointer = '*' in declaration
            bound_pointer = ctype is ctypes.c_void_p or ctype is ctypes.c_char_p or hasattr(ctype, 'contents')
            assert is_pointer == bound_pointer, f"{name}: '{declaration}' bound as {ctype}"
            if not is_pointer and re.match(r'(const\s+)?double\b', declaration):
                assert ctype is ctypes.c_double, f"{name}: '{declaration}' bound as {ctype}"


def test_version_and_strerror(lib):
    assert lib.okp_version() == 2
    assert lib.okp_strerror(0) == b"ok"
    assert b"NULL" in lib.okp_strerror(-1)


def test_struct_sizes_match_header():
    # OkpCamera: 4 + 4 + 9 doubles + 2 int32; OkpDecodeParams: see header; tables: 16 pointers
    assert ctypes.sizeof(_abi.OkpCamera) == 17 * 8 + 8
    assert ctypes.sizeof(_abi.OkpDecodeParams) == 56
    assert _abi.OkpDecodeParams.top_k.offset == 40
    assert ctypes.sizeof(_abi.OkpDecodeTables) == 16 * ctypes.sizeof(ctypes.c_void_p)
    assert _abi.OkpDecodeParams.outlier_distance.offset == 16


def test_argument_validation_happens_on_the_host(lib):
    prm = _abi.make_params()
    tables = _abi.OkpDecodeTables()
    assert lib.okp_decode_workspace_bytes(4, 3, 64, 64, ctypes.byref(prm)) > 0
    assert lib.okp_decode_workspace_bytes(4, 99, 64, 64, ctypes.byref(prm)) == 0
    # NULL heat pointer / bad shapes are rejected before any launch
    assert lib.okp_extract_peaks_f32(None, 1, 3, 64, 64, ctypes.byref(prm), ctypes.byref(tables), None, 0, None) == -1
    assert lib.okp_extract_peaks_f32(None, 1, 0, 64, 64, ctypes.byref(prm), ctypes.byref(tables), None, 0, None) == -2
    assert lib.okp_extract_peaks_f32(None, 0, 3, 64, 64, ctypes.byref(prm), ctypes.byref(tables), None, 0, None) == 0
    bad = _abi.make_params()
    bad.nms_size = 7
    assert lib.okp_extract_peaks_f32(None, 1, 3, 64, 64, ctypes.byref(bad), ctypes.byref(tables), None, 0, None) == -4
    bad = _abi.make_params()
    bad.top_k = 64                                            # more than max_peaks
    assert lib.okp_extract_peaks_f32(None, 1, 3, 64, 64, ctypes.byref(bad), ctypes.byref(tables), None, 0, None) == -3
    bad = _abi.make_params()
    bad.max_peaks = 100000
    assert lib.okp_extract_peaks_f32(None, 1, 3, 64, 64, ctypes.byref(bad), ctypes.byref(tables), None, 0, None) == -3
    assert lib.okp_triangulate_f64(None, None, None, 0, 4, 999, None, None) == -2
    assert lib.okp_triangulate_f64(None, None, None, 0, 4, 2, None, None) == -1


def test_pack_camera_follows_reference_clip_rule():
    from object_keypoints_b200 import synthetic
    cam = synthetic.default_camera((180, 320))
    packed = _abi.pack_camera(cam)
    assert (packed.clip_x, packed.clip_y) == (179, 319)       # pipeline.py:162: (H-1, W-1) applied to (x, y)
    np.testing.assert_allclose(np.array(packed.kinv[:]).reshape(3, 3) @ cam.K, np.eye(3), atol=1e-12)
    cam64 = synthetic.default_camera((64, 64))
    packed = _abi.pack_camera(cam64)
    assert (packed.clip_x, packed.clip_y) == (63, 63)
    np.testing.assert_allclose(cam64.K, [[62.09386781, 0, 32.0896], [0, 62.15028827, 32.72572519], [0, 0, 1]], atol=1e-3)


def test_keypoint_config_forms():
    assert _abi.check_keypoint_config({'keypoint_config': [1, 3]}) == [1, 3]
    assert _abi.check_keypoint_config([1, 1, 1]) == [1, 1, 1]
    with pytest.raises(ValueError):
        _abi.check_keypoint_config([0])
    with pytest.raises(ValueError):
        _abi.check_keypoint_config([1] * 16)


def test_no_cpu_fallback_in_the_product():
    """Nothing under object_keypoints_b200/ may import the oracle."""
    pkg = os.path.join(ROOT, 'object_keypoints_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text and 'liboracle' not in text, f
