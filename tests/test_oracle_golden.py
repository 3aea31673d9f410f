"""Pin the CPU oracle (oracle/*.py) to outputs of the real reference.

The goldens were produced by tests/golden/make_golden.py, which imports the reference's
own Squeeze_ErNET / Squeeze_RedConv (model/squeeze_ernet.py, model/squeeze_ernet_redconv.py)
and the torchvision/Pillow eval transform (dataloaders/aider.py:421-426).
"""
import os

import numpy as np
import pytest

import fixtures
from oracle import ernet_numpy as E
from oracle import ingest_numpy as I

WSETS = ("shipped", "w3", "w3neg")


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_param_and_mac_counts(arch):
    # model_summary/squeeze_ernet.txt:45,48 and squeeze_redconv.txt:48,51
    want_params = {"squeeze-ernet": 169241, "squeeze-redconv": 109569}[arch]
    assert E.count_params(arch) == want_params
    macs = E.count_macs(arch)
    want_macs = {"squeeze-ernet": 45.64e6, "squeeze-redconv": 38.89e6}[arch]   # torchinfo adds bias terms
    assert abs(macs - want_macs) / want_macs < 0.01
    assert [k for k, _ in fixtures.key_shapes(arch)] == list(E.expected_keys(arch).keys()) or \
        set(k for k, _ in fixtures.key_shapes(arch)) == set(E.expected_keys(arch).keys())
    assert len(E.expected_keys(arch)) == {"squeeze-ernet": 56, "squeeze-redconv": 62}[arch]


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_shipped_weights_layout(arch):
    sd = fixtures.shipped_state_dict(arch)
    want = E.expected_keys(arch)
    assert set(sd.keys()) == set(want.keys())
    for k, shp in want.items():
        assert tuple(sd[k].shape) == tuple(shp), k


@pytest.mark.parametrize("arch", fixtures.ARCHS)
@pytest.mark.parametrize("wset", WSETS)
def test_forward_fp64_matches_reference(arch, wset, model_golden):
    sd = fixtures.get_state_dict(arch, wset)
    for iname, x in (("norm", fixtures.normal_tensors(4, seed=7)), ("frames", model_golden["x_frames"])):
        tag = f"{arch}/{wset}/{iname}"
        res = E.forward(sd, x, arch, dtype=np.float64, want_taps=(iname == "norm"))
        ref = model_golden[f"{tag}/logits64"]
        scale = np.abs(ref).max()
        assert np.abs(res["logits"] - ref).max() <= 1e-9 * scale, tag
        assert np.abs(res["probs"] - model_golden[f"{tag}/probs64"]).max() <= 1e-9
        assert (res["probs"].argmax(1) == model_golden[f"{tag}/probs64"].argmax(1)).all()
        if iname == "norm":
            names = {"conv1": None, "stem": None}
            taps = res["taps"]
            # oracle tap name -> reference module name
            mapping = {"acff1": "acff1", "pool1": "pool1", "acff2": "acff2", "pool2": "pool2",
                       "acff3": "acff3", "acff4": "acff4"}
            mapping["stem"] = "conv_red1" if arch == "squeeze-redconv" else "conv1"
            mapping["pool3"] = "conv_red3" if arch == "squeeze-redconv" else "pool3"
            del names
            for mine, theirs in mapping.items():
                sub = model_golden[f"{tag}/tap/{theirs}/sub"]
                got = taps[mine][0, :, ::5, ::5]
                assert got.shape == sub.shape, (mine, got.shape, sub.shape)
                s = max(np.abs(sub).max(), 1e-30)
                assert np.abs(got - sub).max() <= 1e-9 * s, (tag, mine)
                asum = float(model_golden[f"{tag}/tap/{theirs}/abs_sum"])
                assert abs(np.abs(taps[mine]).sum() - asum) <= 1e-9 * asum, (tag, mine)


@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_forward_fp32_close_to_reference_fp32(arch, model_golden):
    sd = fixtures.get_state_dict(arch, "shipped")
    x = fixtures.normal_tensors(4, seed=7)
    res = E.forward(sd, x, arch, dtype=np.float32)
    ref = model_golden[f"{arch}/shipped/norm/logits32"]
    assert np.abs(res["logits"] - ref).max() <= 1e-4 * np.abs(ref).max()


def test_forward_rejects_240():
    # the reference cannot run the squeeze models at 240x240 (view(-1,20) fails / mixes images)
    sd = fixtures.get_state_dict("squeeze-ernet", "w3")
    with pytest.raises(ValueError):
        E.forward(sd, np.zeros((1, 3, 240, 240), np.float32), "squeeze-ernet")
    with pytest.raises(ValueError):
        E.forward(sd, np.zeros((1, 3, 140, 140), np.float32), "ernet")


def _case_frame(name):
    import importlib.util, os
    spec = importlib.util.spec_from_file_location(
        "make_golden_cases", os.path.join(fixtures.GOLDEN, "make_golden.py"))
    # parse INGEST_CASES without importing the reference: read the literal from the source
    src = open(spec.origin).read()
    start = src.index("INGEST_CASES = [")
    end = src.index("]\n", start) + 1
    ns = {}
    exec(src[start:end], ns)
    for n, kind, h, w, seed in ns["INGEST_CASES"]:
        if n == name:
            f = fixtures.noise_frames(1, h, w, seed) if kind == "noise" else fixtures.smooth_frames(1, h, w, seed)
            return f[0]
    raise KeyError(name)


INGEST_NAMES = ["noise240", "smooth240", "noise480x640", "smooth350x372", "noise372x350",
                "noise100x120_up", "noise159x159_id", "smooth720x1280", "noise161x300"]


@pytest.mark.parametrize("name", INGEST_NAMES + ["real240"])
def test_ingest_bit_exact(name, ingest_golden):
    frame = ingest_golden["real240/frame"] if name == "real240" else _case_frame(name)
    crop = I.crop_u8(frame)
    assert np.array_equal(crop, ingest_golden[f"{name}/crop_u8"]), name
    t = I.ingest(frame[None])[0]
    ref = ingest_golden[f"{name}/tensor"]
    assert t.dtype == np.float32 and np.array_equal(t, ref), name


def test_ingest_constants():
    assert I.resized_size(240, 240) == (159, 159)
    assert I.resized_size(480, 640) == (159, 212)
    assert I.center_crop_offset(159) == 10          # round(9.5) -> 10 (banker's)
    xmin, xlen, kk = I.resample_coeffs(240, 159)
    assert kk.shape == (159, 5) and int(kk.sum(1).min()) >= (1 << 22) - 3


# ------------------------------------------------------------------------------------ ErNET
@pytest.mark.parametrize("wset", ["shipped", "w3", "w3neg"])
def test_ernet_oracle_matches_reference_goldens(wset):
    """oracle/ernet_numpy.forward_ernet against outputs of the real reference class (tests/golden/make_golden_ernet.py):
    fc output, probabilities and sub-sampled intermediates of every layer, fp64."""
    import fixtures
    from oracle import ernet_numpy as E
    g = np.load(os.path.join(fixtures.GOLDEN, "ernet_golden.npz"))
    sd = fixtures.get_state_dict("ernet", wset)
    x = fixtures.normal_tensors(3, seed=17, hw=240)
    r = E.forward(sd, x, "ernet", dtype=np.float64, want_taps=True)
    tag = f"ernet/{wset}/norm"
    ref = g[f"{tag}/logits64"]
    assert np.abs(r["logits"] - ref).max() <= 1e-9 * np.abs(ref).max()
    assert np.abs(r["probs"] - g[f"{tag}/probs64"]).max() <= 1e-9
    for name, v in r["taps"].items():
        key = f"{tag}/tap/{'conv1' if name == 'stem' else name}"
        if f"{key}/shape" not in g.files:
            continue                                   # oracle-internal tap without a reference module of that name
        assert tuple(g[f"{key}/shape"]) == v.shape, name
        assert np.abs(v[0, :, ::5, ::5] - g[f"{key}/sub"]).max() <= 1e-9 * max(1.0, np.abs(g[f"{key}/sub"]).max()), name
    r32 = E.forward(sd, x, "ernet", dtype=np.float32)
    assert np.abs(r32["logits"] - g[f"{tag}/logits32"]).max() <= 2e-4 * np.abs(ref).max()


def test_ernet_packer_head_collapse():
    """The packer folds conv2 -> AvgPool(5,1,0) -> view -> fc into W_eff[5][49][256]: check it against the oracle's
    un-collapsed head on random activations, and the blob ids of the six blocks."""
    import fixtures
    import rtdm_b200
    from oracle import ernet_numpy as E
    from rtdm_b200 import pack as P
    sd = fixtures.get_state_dict("ernet", "w3")
    t = P.derive_simt(sd, "ernet")
    assert t[P.T_EHEAD_W].shape == (5, 49, 256)
    for k, (c, co) in enumerate(P.widths("ernet")):
        assert t[P.block_base(k) + P.T_PW_W].shape == (3 * c, co)
    a6 = np.random.RandomState(3).standard_normal((2, 256, 7, 7))
    z = E.avg_pool_5x5_s1_p0(E.conv2d_pointwise(a6, np.asarray(sd["conv2.weight"], np.float64), None))
    want = z.reshape(-1, 45) @ np.asarray(sd["fc.weight"], np.float64).T + np.asarray(sd["fc.bias"], np.float64)
    got = np.einsum("opk,bkp->bo", t[P.T_EHEAD_W], a6.reshape(2, 256, 49)) + t[P.T_HEAD_B]
    assert np.abs(got - want).max() <= 1e-10 * np.abs(want).max()
    blob = rtdm_b200.pack_state_dict(sd, "ernet", "bf16")
    assert len(blob) > 4 * sum(int(np.prod(v.shape)) for v in t.values())
    with pytest.raises(ValueError):
        rtdm_b200.pack_state_dict(sd, "ernet", "int8")


# ------------------------------------------------------------------------------------ ErNET's transform (240 crop)
INGEST240_CASES = [("noise240x240", "noise", 240, 240, 31), ("smooth480x640", "smooth", 480, 640, 32), ("noise300x280", "noise", 300, 280, 33),
                   ("noise273x273_id", "noise", 273, 273, 34), ("smooth260x420", "smooth", 260, 420, 35)]


def ingest240_frame(kind, h, w, seed):
    return (fixtures.noise_frames if kind == "noise" else fixtures.smooth_frames)(1, h, w, seed)[0]


@pytest.fixture(scope="module")
def ingest240_golden():
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ingest240_golden.npz"))


@pytest.mark.parametrize("case", INGEST240_CASES, ids=[c[0] for c in INGEST240_CASES])
def test_ingest240_oracle_bit_exact_with_torchvision(case, ingest240_golden):
    """aider_transforms = Resize(273) -> CenterCrop(240) -> ToTensor -> Normalize (dataloaders/aider.py:412-426,430)."""
    name, kind, h, w, seed = case
    f = ingest240_frame(kind, h, w, seed)
    assert np.array_equal(I.crop_u8(f, 240), ingest240_golden[f"{name}/crop_u8"])
    if name == "noise240x240":
        assert np.array_equal(I.ingest(f[None], 240)[0], ingest240_golden["noise240x240/tensor"])


@pytest.mark.parametrize("arch", fixtures.ARCHS)
@pytest.mark.parametrize("wset", WSETS)
def test_torch_restatement_matches_reference(arch, wset, model_golden):
    """oracle/ernet_torch.forward is the timed CPU arm of bench.py and the fp32 reference of the int8 agreement tests:
    pin it to the outputs of the real reference classes too (logits64 for every weight set; the fp32 run of the shipped
    checkpoint must reproduce the reference's own fp32 logits to fp32 rounding)."""
    import torch
    from oracle import ernet_torch as T
    sd_np = fixtures.get_state_dict(arch, wset)
    for iname, x in (("norm", fixtures.normal_tensors(4, seed=7)), ("frames", model_golden["x_frames"])):
        tag = f"{arch}/{wset}/{iname}"
        ref = model_golden[f"{tag}/logits64"]
        p64, l64 = T.forward(T.to_torch_sd(sd_np, torch.float64), torch.from_numpy(x).double(), arch)
        assert np.abs(l64.numpy() - ref).max() <= 1e-9 * np.abs(ref).max(), tag
        assert np.abs(p64.numpy() - model_golden[f"{tag}/probs64"]).max() <= 1e-9
        p32, l32 = T.forward(T.to_torch_sd(sd_np), torch.from_numpy(x), arch)
        assert l32.dtype == torch.float32
        assert np.abs(l32.double().numpy() - ref).max() <= 2e-5 * np.abs(ref).max(), tag
        assert (p32.argmax(1).numpy() == model_golden[f"{tag}/probs64"].argmax(1)).all()
    ref32 = model_golden[f"{arch}/shipped/norm/logits32"]
    _, l32 = T.forward(T.to_torch_sd(fixtures.get_state_dict(arch, "shipped")), torch.from_numpy(fixtures.normal_tensors(4, seed=7)), arch)
    assert np.abs(l32.numpy() - ref32).max() <= 2e-6 * np.abs(ref32).max()


def test_real_frames_oracle_matches_reference():
    """Input set I3: the 15 AIDER JPEGs of the reference tree, as the 140x140 crops torchvision/Pillow produced
    (tests/golden/make_golden_real.py).  The oracle's normalisation + forward reproduces the real classes' fp64 logits."""
    g = np.load(os.path.join(fixtures.GOLDEN, "real_golden.npz"))
    crops = g["crops_u8"]
    assert crops.shape == (15, 140, 140, 3) and crops.dtype == np.uint8
    lut = I.normalise_lut()
    x = np.stack([np.stack([lut[c[:, :, ch], ch] for ch in range(3)], 0) for c in crops], 0)
    for arch in fixtures.ARCHS:
        res = E.forward(fixtures.get_state_dict(arch, "shipped"), x, arch, dtype=np.float64)
        ref = g[f"{arch}/shipped/logits64"]
        assert np.abs(res["logits"] - ref).max() <= 1e-6 * np.abs(ref).max(), arch      # fp32 normalised input on both sides
        assert (res["probs"].argmax(1) == g[f"{arch}/shipped/probs64"].argmax(1)).all()
