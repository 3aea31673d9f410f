"""No-GPU checks of the C-ABI library: it loads, exports every symbol include/ernet_b200.h declares,
fails loudly without a device, and its host-side ingest tables are bit-identical to the oracle's."""
import ctypes as C
import os
import re

import numpy as np
import pytest
import torch

import fixtures
import rtdm_b200
from oracle import ingest_numpy as I
from rtdm_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HAS_GPU = torch.cuda.is_available()


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    hdr = open(os.path.join(ROOT, "include", "ernet_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(ernet_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 17
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
        assert name in _lib.SIGNATURES, f"{name} has no ctypes signature"
    assert lib.ernet_abi_version() == 1


@pytest.mark.skipif(HAS_GPU, reason="checks the no-device error path")
def test_create_fails_loudly_without_gpu():
    lib = _lib.load()
    h = C.c_void_p()
    rc = lib.ernet_create(C.byref(h), 0, 0, 0)
    assert rc < 0 and not h.value
    assert len(_lib.last_error()) > 0
    with pytest.raises(RuntimeError):
        _lib.check(rc)


def test_create_rejects_bad_enums_before_touching_cuda():
    lib = _lib.load()
    h = C.c_void_p()
    assert lib.ernet_create(C.byref(h), 7, 0, 0) == _lib.ERR_INVALID_ARG
    assert "Unsupported model" in _lib.last_error()
    assert lib.ernet_create(C.byref(h), 0, 9, 0) == _lib.ERR_INVALID_ARG
    assert lib.ernet_create(None, 0, 0, 0) == _lib.ERR_INVALID_ARG


@pytest.mark.parametrize("hw", [(240, 240), (480, 640), (350, 372), (372, 350), (100, 120), (159, 159),
                                (720, 1280), (161, 300), (1080, 1920)])
def test_ingest_tables_bit_identical_to_oracle(hw):
    H, W = hw
    lib = _lib.load()
    meta = np.zeros(8, np.int32)
    xmin, xlen, ymin, ylen = (np.zeros(140, np.int32) for _ in range(4))
    kx = np.zeros(140 * 64, np.int32)
    ky = np.zeros(140 * 64, np.int32)
    lut = np.zeros(768, np.float32)
    p = lambda a: a.ctypes.data
    _lib.check(lib.ernet_ingest_tables_host(H, W, p(meta), p(xmin), p(xlen), p(kx), p(ymin), p(ylen), p(ky), p(lut)))
    nh, nw = I.resized_size(H, W)
    top, left = I.center_crop_offset(nh), I.center_crop_offset(nw)
    assert tuple(meta[:4]) == (nh, nw, top, left)
    assert np.array_equal(lut.reshape(256, 3), I.normalise_lut())
    for size_in, size_out, first, mn, ln, kk, ks in ((W, nw, left, xmin, xlen, kx, meta[5]), (H, nh, top, ymin, ylen, ky, meta[4])):
        if size_in == size_out:
            assert ks == 1 and np.array_equal(mn, np.arange(first, first + 140)) and (ln == 1).all()
            assert (kk[:140] == 1 << 22).all()
            continue
        omn, oln, okk = I.resample_coeffs(size_in, size_out)
        assert ks == okk.shape[1]
        assert np.array_equal(mn, omn[first:first + 140])
        assert np.array_equal(ln, oln[first:first + 140])
        assert np.array_equal(kk[:140 * ks].reshape(140, ks), okk[first:first + 140])


def test_ingest_tables_reject_tiny_frames():
    lib = _lib.load()
    z = np.zeros(140 * 64, np.int32)
    rc = lib.ernet_ingest_tables_host(0, 10, z.ctypes.data, z.ctypes.data, z.ctypes.data, z.ctypes.data,
                                      z.ctypes.data, z.ctypes.data, z.ctypes.data, None)
    assert rc == _lib.ERR_INVALID_ARG


# ---------------------------------------------------------------- model surface (host logic)
@pytest.mark.parametrize("arch", fixtures.ARCHS)
def test_shim_state_dict_surface(arch):
    cls = {"squeeze-ernet": rtdm_b200.Squeeze_ErNET, "squeeze-redconv": rtdm_b200.Squeeze_RedConv}[arch]
    m = cls()                                           # zero-arg constructor like the reference
    sd = m.state_dict()
    want = fixtures.key_shapes(arch)
    assert list(sd.keys()) == [k for k, _ in want]      # same keys, same registration order
    for k, shp in want:
        assert tuple(sd[k].shape) == tuple(shp), k
    n_params = sum(p.numel() for p in m.parameters())
    assert n_params == {"squeeze-ernet": 169241, "squeeze-redconv": 109569}[arch]   # model_summary/*.txt
    shipped = {k: torch.from_numpy(np.asarray(v)) for k, v in fixtures.shipped_state_dict(arch).items()}
    m.load_state_dict(shipped)                          # strict
    assert next(m.parameters()).dtype == torch.float32
    assert m.half().precision == "fp16" and next(m.parameters()).dtype == torch.float16
    seen = []
    m.apply(lambda mod: seen.append(type(mod).__name__))
    assert "Conv2d" in seen and "BatchNorm2d" in seen and "Linear" in seen


def test_shim_errors_without_gpu_or_in_train_mode():
    m = rtdm_b200.Squeeze_ErNET()
    x = torch.zeros(1, 3, 140, 140)
    with pytest.raises(ValueError):
        m.eval()(torch.zeros(1, 3, 240, 240))           # the reference cannot run 240x240 either
    if not HAS_GPU:
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            m.eval()(x)
    with pytest.raises(ValueError):
        rtdm_b200.Squeeze_ErNET(precision="fp64")
    with pytest.raises(ValueError, match="Unsupported model"):
        rtdm_b200.load_model("yolov3", "/nonexistent", "cpu")       # "ernet" is a supported name since the ErNET path exists


@pytest.mark.parametrize("arch", ["squeeze-ernet", "squeeze-redconv", "ernet"])
@pytest.mark.parametrize("wrapped", [False, True])
def test_load_model_reads_real_checkpoints(arch, wrapped, tmp_path):
    """load_model success path (aider-predict.py:35-41): a plain state_dict .pt file and a training checkpoint of the
    form {'model_state_dict': ..., 'epoch': ...} both load, strictly, into an eval-mode model on the asked device."""
    sd = {k: torch.from_numpy(np.asarray(v)) for k, v in fixtures.shipped_state_dict(arch).items()}
    path = tmp_path / f"{arch}.pt"
    torch.save({"model_state_dict": sd, "epoch": 45, "optimizer_state_dict": {}} if wrapped else sd, path)
    m = rtdm_b200.load_model(arch, str(path), "cpu")
    assert not m.training and type(m).ARCH == arch
    got = m.state_dict()
    assert list(got.keys()) == list(sd.keys())
    for k in sd:
        assert torch.equal(got[k], sd[k]), k
    with pytest.raises(RuntimeError):                   # a checkpoint of another architecture must not load silently
        other = "squeeze-redconv" if arch != "squeeze-redconv" else "squeeze-ernet"
        rtdm_b200.load_model(other, str(path), "cpu")


def test_bench_configuration_rules(monkeypatch):
    """bench.py: N=1 runs BASELINE configs[1] (256 frames), N>1 runs configs[4] (8192 frames split N ways, strong
    scaling) unless --batch pins the per-GPU batch; both arms describe the configuration with the same `config` object."""
    import argparse
    import bench
    a = argparse.Namespace(batch=None, arch="squeeze-ernet", precision="bf16")
    assert bench.resolve_batch(a, 1) == (256, None)
    assert bench.resolve_batch(a, 2) == (4096, 8192) and bench.resolve_batch(a, 8) == (1024, 8192)
    a.batch = 512
    assert bench.resolve_batch(a, 4) == (512, None)
    c1 = bench.common_config("squeeze-ernet", "bf16", 1024, 8, 8192)
    assert "configs[4]" in c1["workload"] and c1["global_batch"] == 8192
    assert "configs[1]" in bench.common_config("squeeze-ernet", "bf16", 256, 1, None)["workload"]
    # block 1: 36 regular tiles + the 2 tiles of the tail unit per image (45 regular tiles with ERNET_TAIL_TILES=0)
    monkeypatch.delenv("ERNET_TAIL_TILES", raising=False)
    ops = bench.issued_flops("squeeze-ernet", "int8")
    assert ops["tc_block1"] == 38 * 13 * 2 * 128 * 64 * 32 and ops["tc_block2"] == 8 * 25 * 2 * 2 * 128 * 96 * 32
    assert bench.issued_flops("squeeze-ernet", "bf16")["tc_block1"] == 38 * 25 * 2 * 128 * 64 * 16
    monkeypatch.setenv("ERNET_TAIL_TILES", "0")
    assert bench.issued_flops("squeeze-ernet", "bf16")["tc_block1"] == 45 * 25 * 2 * 128 * 64 * 16


def test_build_detects_changed_sources(tmp_path):
    """build(): the library carries the sha256 of the sources it was compiled from; a changed source forces a rebuild
    (file times are not trusted)."""
    import importlib
    b = importlib.import_module("real-time-disaster-management_b200.build")
    assert b.embedded_hash() == b.source_hash() and not b.needs_build()
    assert _lib.load().ernet_source_hash().decode() == b.source_hash()
    assert b.embedded_hash(str(tmp_path / "missing.so")) is None
