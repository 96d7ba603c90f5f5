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
    # the three precision modes: fast < mixed (3xTF32 forward, no lo halves in the backward) < parity; eval needs far less
    by_mode = {}
    for mode in (1, 2, 3):
        assert lib.tf_model_workspace_bytes(h, 8, 960, 1280, 1, mode, ctypes.byref(sz)) == 0
        by_mode[mode] = sz.value
    assert by_mode[1] < by_mode[3] < by_mode[2] < 60 * 2 ** 30
    assert lib.tf_model_workspace_bytes(h, 8, 960, 1280, 0, 3, ctypes.byref(sz)) == 0
    ev3 = sz.value
    assert lib.tf_model_workspace_bytes(h, 8, 960, 1280, 0, 2, ctypes.byref(sz)) == 0
    assert ev3 == sz.value < by_mode[1]                                                 # inference: mixed == parity
    assert lib.tf_model_workspace_bytes(h, 8, 960, 1280, 1, 4, ctypes.byref(sz)) != 0     # unknown mode
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


# ------------------------------------------------------------------------------------------- GEMM launch planning (host)
def _plan(B, H, W, Cin, Cout, k, stride=1, nseg=1, plain=1, has_res=0, sms=148):
    from tinyfaces_b200 import _lib
    out = (ctypes.c_int * 16)()
    assert _lib.lib().tf_conv_plan(B, H, W, Cin, Cout, k, stride, nseg, plain, has_res, sms, out) == 0
    keys = ["spatial", "tw", "th", "tiles_x", "tiles_y", "m_tiles", "bn", "n_tiles", "grid", "two_cta", "main_tiles", "ksplit",
            "kiters", "tail_lo", "tail_hi"]
    d = dict(zip(keys, list(out)))
    d["tail_pix0"] = -1 if d["tail_lo"] < 0 else (d["tail_hi"] << 31) | d["tail_lo"]
    return d


def test_conv_plan_known_answers_and_invariants():
    """tf_conv_plan (pure host arithmetic): the bench's layer-3 shapes, and invariants over a sweep of shapes."""
    p = _plan(8, 60, 80, 256, 256, 3)                     # layer-3 3x3 at batch-8 960x1280: 320 tiles on 148 SMs
    assert (p["bn"], p["m_tiles"], p["grid"], p["two_cta"]) == (256, 320, 148, 0)
    assert (p["main_tiles"], p["ksplit"], p["kiters"]) == (295, 5, 72)              # 25 tail tiles cut into 5 K slices
    assert p["tail_pix0"] == (7 * 60 + 24) * 80           # image 7, tile row 3 (th = 8): the split rows are contiguous to the end
    p = _plan(8, 60, 80, 1024, 256, 1)                    # layer-3 conv1: flat, K-heavy, one n-tile -> 2-CTA tiles
    assert (p["spatial"], p["bn"], p["two_cta"], p["ksplit"]) == (0, 256, 1, 1)
    p = _plan(8, 60, 80, 256, 1024, 1)                    # layer-3 conv3: 1200 tiles, no split (1x1), one-CTA kernel
    assert (p["bn"], p["n_tiles"], p["grid"], p["two_cta"], p["ksplit"], p["main_tiles"]) == (256, 4, 148, 0, 1, 1200)
    assert _plan(8, 60, 80, 256, 256, 3, plain=0)["ksplit"] == 1          # a fused epilogue cannot take K slices
    assert _plan(8, 60, 80, 1024, 256, 1, has_res=1)["two_cta"] == 0      # the residual epilogue lives in the one-CTA kernel
    r = np.random.RandomState(0)
    for _ in range(300):
        B, H, W = int(r.randint(1, 9)), int(r.randint(4, 200)), int(r.randint(4, 200))
        Cin, Cout = int(r.choice([32, 64, 128, 256, 512, 1024])), int(r.choice([64, 128, 256, 512, 1024]))
        k, stride, nseg = int(r.choice([1, 3])), int(r.choice([1, 2])), int(r.choice([1, 3]))
        p = _plan(B, H, W, Cin, Cout, k, stride, nseg)
        Ho, Wo = ((H + 1) // 2, (W + 1) // 2) if stride == 2 else (H, W)
        tiles = p["m_tiles"] * p["n_tiles"]
        assert p["bn"] in (64, 128, 256) and Cout % p["bn"] == 0 and p["n_tiles"] == Cout // p["bn"]
        assert 0 < p["grid"] <= 148 and p["grid"] % p["n_tiles"] == 0                 # one n-tile per CTA (BN statistics rows)
        assert p["kiters"] == nseg * k * k * Cin // 32
        if p["spatial"]:
            assert p["tw"] * p["th"] == 128 and p["tiles_x"] * p["tw"] >= Wo and p["tiles_y"] * p["th"] >= Ho
            assert p["m_tiles"] == B * p["tiles_x"] * p["tiles_y"]
        else:
            assert p["m_tiles"] == (B * Ho * Wo + 127) // 128
        if p["ksplit"] > 1:
            assert k == 3 and not p["two_cta"] and 0 < p["main_tiles"] < tiles and p["main_tiles"] % p["n_tiles"] == 0
            tail = tiles - p["main_tiles"]
            assert tail * p["ksplit"] <= p["grid"] and p["ksplit"] * 8 <= p["kiters"]      # one round of K slices, >= 8 K steps each
            main_m = p["main_tiles"] // p["n_tiles"]
            assert main_m % p["tiles_x"] == 0                                            # split region starts at a tile-row boundary
            img, ty = divmod(main_m, p["tiles_x"] * p["tiles_y"])[0], (main_m % (p["tiles_x"] * p["tiles_y"])) // p["tiles_x"]
            assert p["tail_pix0"] == (img * Ho + ty * p["th"]) * Wo and 0 <= p["tail_pix0"] < B * Ho * Wo
        else:
            assert p["main_tiles"] == tiles and p["tail_pix0"] == -1


def test_stride2_dgrad_parity_class_taps_reproduce_autograd():
    """The tap tables of conv_dgrad_s2 (tf_dgrad_s2_taps), evaluated with numpy on the CPU, give the input gradient of a
    stride-2 convolution: dx[2i+py, 2j+px] = sum_t dy[i+oy_t, j+ox_t] . Wflip[tap_t]  (no zero insertion)."""
    import torch
    from tinyfaces_b200 import _lib
    lib = _lib.lib()
    gen = torch.Generator().manual_seed(0)
    for (k, H, W) in [(3, 9, 12), (3, 10, 7), (3, 8, 8), (1, 9, 12), (1, 6, 5)]:
        Ci, Co, B = 5, 4, 2
        x = torch.zeros(B, Ci, H, W, dtype=torch.float64, requires_grad=True)
        w = torch.randn(Co, Ci, k, k, generator=gen, dtype=torch.float64)
        y = torch.nn.functional.conv2d(x, w, stride=2, padding=k // 2)
        dy = torch.randn(y.shape, generator=gen, dtype=torch.float64)
        y.backward(dy)
        ref = x.grad.numpy()                                             # [B, Ci, H, W]
        Ho, Wo = y.shape[2], y.shape[3]
        wflip = w.flip(2, 3).permute(1, 2, 3, 0).reshape(Ci, k * k, Co).numpy()      # the packed dgrad layout [ci][flipped tap][co]
        dyp = np.zeros((B, Co, Ho + 1, Wo + 1))
        dyp[:, :, :Ho, :Wo] = dy.numpy()                                 # reads past the edge are zero (TMA out-of-bounds fill)
        got = np.zeros_like(ref)
        ntot = 0
        for py in range(2):
            for px in range(2):
                tw_, ox, oy = (ctypes.c_int * 9)(), (ctypes.c_int * 9)(), (ctypes.c_int * 9)()
                nt = lib.tf_dgrad_s2_taps(k, py, px, tw_, ox, oy)
                ntot += nt
                Hq, Wq = (H - py + 1) // 2, (W - px + 1) // 2
                for t in range(nt):
                    patch = dyp[:, :, oy[t]:oy[t] + Hq, ox[t]:ox[t] + Wq]                  # [B, Co, Hq, Wq]
                    got[:, :, py::2, px::2] += np.einsum("bohw,io->bihw", patch, wflip[:, tw_[t], :])
        assert ntot == (9 if k == 3 else 1)
        np.testing.assert_allclose(got, ref, rtol=1e-12, atol=1e-12)
    assert lib.tf_dgrad_s2_taps(2, 0, 0, tw_, ox, oy) == -1


def test_package_synthetic_weights_equal_the_test_side_recipe():
    """bench.py scores its forward against tests/golden/cfg2_b8_fwd.npz using tinyfaces_b200.synthetic.state_dict (the
    product may not import oracle/): the two generators must agree draw for draw."""
    import torch
    from oracle import synth
    from tinyfaces_b200 import synthetic
    a = synth.synthetic_state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    b = synthetic.state_dict(seed=1, bn3_gamma=0.25, beta_jitter=0.1)
    assert set(a) == set(b) and all(torch.equal(a[k], b[k]) for k in a)
