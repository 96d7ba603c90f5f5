"""CPU: the C-ABI library builds/loads and exports exactly what include/tinyfaces_b200.h declares;
host-side logic (sampler RNG protocol, template masks, workspace planner) without touching a GPU."""
import ctypes
import os
import re

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_symbols():
    with open(os.path.join(ROOT, "include", "tinyfaces_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(tf_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from tinyfaces_b200 import _lib
    lib = _lib.lib()
    syms = _header_symbols()
    assert len(syms) >= 20
    for s in syms:
        assert hasattr(lib, s), "missing export " + s
    assert set(_lib.exported_symbols()) <= set(syms) | {"tf_last_error_string"}
    assert lib.tf_version() >= 100


def test_model_param_table_matches_reference_state_dict_names():
    from oracle import synth
    from tinyfaces_b200 import _lib
    lib = _lib.lib()
    h = ctypes.c_void_p()
    assert lib.tf_model_create(25, ctypes.byref(h)) == 0
    names = [lib.tf_model_param_name(h, i).decode() for i in range(lib.tf_model_num_params(h))]
    sd = synth.synthetic_state_dict(seed=0)
    assert set(names) <= set(sd.keys())
    missing = set(sd.keys()) - set(names)
    assert all(k.startswith("model.fc.") or k.endswith("num_batches_tracked") for k in missing)
    # workspace planner (dry run of forward/backward) works without a device
    sz = ctypes.c_size_t()
    assert lib.tf_model_workspace_bytes(h, 8, 960, 1280, 1, 1, ctypes.byref(sz)) == 0
    assert 10 * 2 ** 30 < sz.value < 60 * 2 ** 30
    h3, w3 = ctypes.c_int(), ctypes.c_int()
    for (H, W, e) in [(500, 500, (63, 63)), (960, 1280, (120, 160)), (5000, 5000, (625, 625)), (1250, 1250, (157, 157))]:
        assert lib.tf_model_output_shape(h, H, W, ctypes.byref(h3), ctypes.byref(w3)) == 0
        assert (h3.value, w3.value) == e
    assert lib.tf_model_workspace_bytes(h, 0, 960, 1280, 1, 1, ctypes.byref(sz)) != 0     # bad args fail loudly
    assert b"bad args" in lib.tf_last_error_string()
    lib.tf_model_destroy(h)


def test_host_sampler_consumes_numpy_rng_like_the_oracle():
    from oracle import loss_oracle
    from tinyfaces_b200.models.utils import balance_sampling
    r = np.random.RandomState(3)
    lab = r.choice([-1.0, 0.0, 1.0], size=(25, 12, 16), p=[0.6, 0.2, 0.2]).astype(np.float32)
    a, b = lab.copy(), lab.copy()
    np.random.seed(5)
    balance_sampling(a, 0.5)
    s1 = np.random.get_state()[1][:4].copy()
    np.random.seed(5)
    loss_oracle.balance_sampling(b, 0.5)
    s2 = np.random.get_state()[1][:4].copy()
    assert np.array_equal(a, b) and np.array_equal(s1, s2)
    assert (a == 1).sum() == 128 and (a == -1).sum() == 128


def test_invalid_template_ids_match_oracle():
    from oracle import decode_oracle, synth
    from tinyfaces_b200.models.utils import invalid_template_ids
    t = synth.load_templates()
    for s in (0.25, 0.5, 1, 2, 4, 0.7):
        assert np.array_equal(np.sort(invalid_template_ids(t, s)), np.sort(decode_oracle.invalid_ids(t, s)))


def test_no_cpu_fallback():
    import pytest
    import torch
    from tinyfaces_b200 import ops
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.nms_device(torch.zeros(3, 4, dtype=torch.float64), torch.zeros(3, dtype=torch.float64), 0.3)


def test_pil_resize_tables_are_bit_exact():
    """The host-side restatement of Pillow's bilinear resampling (coefficient tables + two integer passes) that
    feeds tf_pyramid_level matches the real PIL call the reference makes (evaluation.py:46-47) bit for bit."""
    from PIL import Image
    from torchvision import transforms
    from tinyfaces_b200.pyramid import PRECISION_BITS, bilinear_coeffs, resized_size

    def emulate(arr, size):
        H, W, _ = arr.shape
        Wo, Ho = resized_size(W, H, size)
        a = arr.astype(np.int64)
        if Wo != W:
            b, k, _ = bilinear_coeffs(W, Wo)
            out = np.zeros((H, Wo, 3), np.int64)
            for xx in range(Wo):
                x0, c = b[xx]
                ss = (1 << (PRECISION_BITS - 1)) + (a[:, x0:x0 + c, :] * k[xx, :c][None, :, None]).sum(1)
                out[:, xx, :] = np.clip(ss >> PRECISION_BITS, 0, 255)
            a = out
        if Ho != H:
            b, k, _ = bilinear_coeffs(H, Ho)
            out = np.zeros((Ho, a.shape[1], 3), np.int64)
            for yy in range(Ho):
                y0, c = b[yy]
                ss = (1 << (PRECISION_BITS - 1)) + (a[y0:y0 + c] * k[yy, :c][:, None, None]).sum(0)
                out[yy] = np.clip(ss >> PRECISION_BITS, 0, 255)
            a = out
        return a.astype(np.uint8)

    r = np.random.RandomState(0)
    for (H, W, size) in [(50, 64, 25), (50, 64, 100), (37, 53, 12), (64, 50, 200), (31, 31, 44), (40, 60, 40), (200, 216, 282)]:
        arr = r.randint(0, 256, (H, W, 3)).astype(np.uint8)
        ref = np.asarray(transforms.functional.resize(Image.fromarray(arr), size))
        assert np.array_equal(ref, emulate(arr, size)), (H, W, size)
