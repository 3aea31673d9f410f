"""the batched evaluation loop with device ingest, the engine build / export tooling and
the add-fusion ACFF block of the detector half.  CPU tests pin the oracles to the goldens and cover the host logic;
the GPU tests (-m gpu) go through the C ABI and compare with the oracles."""
import argparse
import os

import numpy as np
import pytest
import torch

import fixtures
import rtdm_b200
from oracle import acff_add_numpy as A
from oracle import ernet_numpy as E
from oracle import ingest_numpy as I
from oracle import metrics_numpy as M
from rtdm_b200 import _lib, build_engine as BE, evaluate as EV

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ADD_CASES = ["c128_104", "c256_52", "c128_head14", "c64", "c40_generic"]


@pytest.fixture(scope="module")
def add_golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "acff_add_golden.npz"))


def _add_case(g, name):
    sd = {k.split("/sd/")[1]: g[k] for k in g.files if k.startswith(name + "/sd/")}
    return sd, g[name + "/x"], g[name + "/y64"], g[name + "/y32"]


@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    return torch.device("cuda:0")


# ------------------------------------------------------------------------------------ oracles vs goldens (CPU)
@pytest.mark.parametrize("name", ADD_CASES)
def test_acff_add_oracle_matches_reference_class(name, add_golden):
    sd, x, y64, y32 = _add_case(add_golden, name)
    got = A.acff_add(x.astype(np.float64), sd)
    assert got.shape == y64.shape
    assert np.abs(got - y64).max() <= 1e-12 * np.abs(y64).max()
    assert np.abs(A.acff_add(x, sd) - y32).max() <= 5e-6 * np.abs(y32).max()


def test_metrics_oracle_known_answers():
    pred = [0, 1, 1, 3, 4, 4, 2, 3, 3, 0]
    targ = [0, 1, 2, 3, 4, 3, 2, 3, 0, 0]
    cm = M.confusion_matrix(pred, targ)
    assert cm.sum() == 10 and cm[2, 1] == 1 and cm[3, 4] == 1 and cm[0, 3] == 1 and np.trace(cm) == 7
    mm = M.micro_metrics(cm)
    assert mm["accuracy"] == pytest.approx(0.7) and mm["f1_score"] == mm["precision"] == mm["recall"] == mm["accuracy"]
    pc = M.per_class_metrics(cm)
    assert pc["fire_precision"] == pytest.approx(0.5) and pc["fire_recall"] == pytest.approx(1.0)
    assert pc["normal_precision"] == pytest.approx(2 / 3) and pc["normal_recall"] == pytest.approx(2 / 3)
    assert pc["collapsed building_f1"] == pytest.approx(2 * 1.0 * (2 / 3) / (1.0 + 2 / 3))
    # the product's mirror of compute_per_class_metrics agrees with the restatement, zero rows/columns included
    cm2 = cm.copy()
    cm2[1, :] = 0
    cm2[:, 1] = 0
    for c in (cm, cm2):
        got = EV.compute_per_class_metrics(torch.from_numpy(c))
        want = M.per_class_metrics(c)
        assert set(got) == set(want) and all(got[k] == pytest.approx(want[k]) for k in want)


# ------------------------------------------------------------------------------------ host logic (CPU)
def _write_images(tmp_path, n=23, seed=5):
    from PIL import Image
    rs = np.random.RandomState(seed)
    sizes = [(240, 240), (200, 320), (240, 240), (300, 260)]
    rows, frames = [], []
    for i in range(n):
        h, w = sizes[i % len(sizes)]
        fr = fixtures.smooth_frames(1, h, w, seed=seed * 100 + i)[0] if i % 3 else fixtures.noise_frames(1, h, w, seed=seed * 100 + i)[0]
        cls = int(rs.randint(0, 5))
        rel = f"class{cls}/img{i:03d}.png"
        os.makedirs(os.path.join(tmp_path, f"class{cls}"), exist_ok=True)
        Image.fromarray(fr).save(os.path.join(tmp_path, rel))
        rows.append((rel, cls))
        frames.append(fr)
    rows.append(("missing/nope.jpg", 3))                                   # unreadable file -> blank 240x240 frame (aider.py:53-56)
    frames.append(np.zeros((240, 240, 3), np.uint8))
    csv_path = os.path.join(tmp_path, "split.csv")
    with open(csv_path, "w") as f:
        for rel, cls in rows:
            f.write(f"{rel},{cls}\n")
    return csv_path, rows, frames


def test_frame_batches_yield_every_sample_once(tmp_path):
    csv_path, rows, frames = _write_images(str(tmp_path))
    samples = EV.read_split(csv_path, str(tmp_path))
    assert [lbl for _, lbl in samples] == [c for _, c in rows]
    seen = []
    for fb, tb in EV.frame_batches(samples, batch_size=4, num_workers=3, pin_memory=False):
        assert fb.dtype == torch.uint8 and fb.dim() == 4 and fb.shape[3] == 3 and 1 <= fb.shape[0] <= 4
        assert tb.dtype == torch.int64 and tb.shape == (fb.shape[0],)
        for fr, t in zip(fb.numpy(), tb.tolist()):
            seen.append((fr.shape, int(fr.astype(np.int64).sum()), t))
    want = sorted((fr.shape, int(fr.astype(np.int64).sum()), c) for fr, (_, c) in zip(frames, rows))
    assert sorted(seen) == want
    with pytest.raises(FileNotFoundError):
        EV.read_split(os.path.join(str(tmp_path), "nope.csv"), str(tmp_path))


def test_trtmodule_state_dict_round_trip_and_errors(tmp_path):
    sd = fixtures.get_state_dict("squeeze-redconv", "w3")
    blob = rtdm_b200.pack_state_dict(sd, "squeeze-redconv", "fp16")
    eng = BE.TRTModule()
    with pytest.raises(RuntimeError):
        eng.state_dict()
    eng.load_state_dict({"format": BE.FORMAT, "arch": "squeeze-redconv", "precision": "fp16",
                         "blob": torch.frombuffer(bytearray(blob), dtype=torch.uint8), "act_scales": None})
    path = os.path.join(str(tmp_path), "squeeze-redconv_fp16_trt.pth")
    torch.save(eng.state_dict(), path)
    back = BE.TRTModule().load_state_dict(torch.load(path, weights_only=True))
    assert back.ARCH == "squeeze-redconv" and back.precision == "fp16" and back.IN_HW == 140
    assert bytes(back.blob.numpy().tobytes()) == blob
    assert next(back.parameters()).dtype == torch.float32                   # callers probe this, build_tensorrt_model.py:162
    with pytest.raises(ValueError):
        BE.TRTModule().load_state_dict({"conv1.weight": torch.zeros(1)})
    with pytest.raises(ValueError):
        BE.TRTModule().load_state_dict({"format": BE.FORMAT, "arch": "resnet", "precision": "fp16", "blob": back.blob})
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            back(torch.zeros(1, 3, 140, 140))                                # no CPU fallback
    assert BE.get_model_input_shape("ernet") == (1, 3, 240, 240) and BE.get_model_input_shape("squeeze-ernet") == (1, 3, 140, 140)
    with pytest.raises(ValueError):
        BE.get_model_input_shape("vgg")


def test_acff_add_module_surface():
    m = rtdm_b200.ACFF(128, 14, 3)
    keys = list(m.state_dict().keys())
    assert keys == ['conv1.weight', 'conv1.bias', 'conv2.weight', 'conv2.bias', 'conv3.weight', 'conv3.bias',
                    'fused_conv.weight', 'fused_conv.bias', 'batch_norm.weight', 'batch_norm.bias',
                    'batch_norm.running_mean', 'batch_norm.running_var', 'batch_norm.num_batches_tracked']
    assert tuple(m.conv2.weight.shape) == (128, 1, 3, 3) and tuple(m.fused_conv.weight.shape) == (14, 128, 1, 1)
    with pytest.raises(RuntimeError):
        m(torch.zeros(1, 128, 8, 8))                                          # train mode
    m.eval()
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError):
            m(torch.zeros(1, 128, 8, 8))                                      # no CPU fallback
    with pytest.raises(ValueError):
        rtdm_b200.ACFF(128, 128, 5)


# ------------------------------------------------------------------------------------ GPU: add-fusion ACFF
@pytest.mark.gpu
@pytest.mark.parametrize("name", ADD_CASES)
def test_acff_add_block_matches_reference(name, add_golden, dev):
    sd, x, y64, _ = _add_case(add_golden, name)
    cin, cout = x.shape[1], y64.shape[1]
    m = rtdm_b200.ACFF(cin, cout, 3)
    m.load_state_dict({k: torch.from_numpy(np.asarray(v)) for k, v in sd.items()})
    m = m.to(dev).eval()
    xt = torch.from_numpy(x).to(dev)
    got = m(xt)
    assert tuple(got.shape) == y64.shape
    err = np.abs(got.double().cpu().numpy() - y64).max() / np.abs(y64).max()
    assert err <= 1e-4, (name, err)                                          # north_star fp32 bound
    # channels_last input takes the no-copy path and gives the same bits
    got2 = m(xt.contiguous(memory_format=torch.channels_last))
    assert torch.equal(got, got2)
    # the depthwise-sum stage alone against the oracle
    lib = _lib.load()
    dw, db = m._pack()[:2]
    B, C, H, W = x.shape
    s = torch.empty((B, H - 2, W - 2, C), dtype=torch.float32, device=dev)
    _lib.check(lib.ernet_acff_add_depthwise(xt.permute(0, 2, 3, 1).contiguous().data_ptr(), 0, B, H, W, C, H - 2, W - 2,
                                            dw.data_ptr(), db.data_ptr(), s.data_ptr(), torch.cuda.current_stream().cuda_stream))
    ref = A.acff_add_depthwise(x.astype(np.float64), {k: np.asarray(v, np.float64) for k, v in sd.items() if "conv" in k})
    assert np.abs(s.permute(0, 3, 1, 2).double().cpu().numpy() - ref).max() <= 2e-6 * np.abs(ref).max()


@pytest.mark.gpu
def test_acff_add_detector_size_linearity(dev):
    """At the detector's real size (yolov3-acffx.cfg:91: 128 channels on a 104x104 map) the oracle is too slow for a
    dense check; the depthwise-sum stage is linear in x for zero biases, so f(a*x1 + x2) == a*f(x1) + f(x2)."""
    B, C, H = 4, 128, 104
    g = torch.Generator(device="cpu").manual_seed(3)
    w = (torch.randn(3, 9, C, generator=g) * 0.3).to(dev)
    b0 = torch.zeros(3, C, device=dev)
    x1 = torch.randn(B, H, H, C, generator=g).to(dev)
    x2 = torch.randn(B, H, H, C, generator=g).to(dev)
    lib = _lib.load()

    def f(x):
        o = torch.empty((B, H - 2, H - 2, C), dtype=torch.float32, device=dev)
        _lib.check(lib.ernet_acff_add_depthwise(x.contiguous().data_ptr(), 0, B, H, H, C, H - 2, H - 2, w.data_ptr(), b0.data_ptr(),
                                                o.data_ptr(), torch.cuda.current_stream().cuda_stream))
        return o
    lhs, rhs = f(2.5 * x1 + x2), 2.5 * f(x1) + f(x2)
    assert float((lhs - rhs).abs().max() / rhs.abs().max()) <= 1e-5
    # and against torch's own conv on the GPU (fp32, TF32 off) as a second opinion at full size
    torch.backends.cudnn.allow_tf32 = False
    xn = x1.permute(0, 3, 1, 2).contiguous()
    ref = sum(torch.nn.functional.conv2d(xn, w[d].t().reshape(C, 1, 3, 3).contiguous(), None, 1, d, d + 1, C) for d in range(3))
    assert float((f(x1).permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max()) <= 1e-5


# ------------------------------------------------------------------------------------ GPU: evaluation loop
@pytest.mark.gpu
def test_confusion_update_matches_oracle(dev):
    rs = np.random.RandomState(11)
    conf = EV.DeviceConfusion(dev)
    preds, targs = [], []
    for n in (1, 7, 256, 1000):
        scores = rs.standard_normal((n, 5)).astype(np.float32)
        scores[::5, 1] = scores[::5, 3] = 9.0                               # ties -> lowest index (torch.argmax on CPU too)
        if n > 5:
            scores[3, 0] = np.nan                                            # a NaN wins (first one), as in torch.argmax
        t = rs.randint(0, 5, n)
        p = conf.update(torch.from_numpy(scores).to(dev), torch.from_numpy(t), want_pred=True).cpu().numpy()
        ref_p = torch.from_numpy(scores).argmax(1).numpy()          # evaluate-classification-metrics.py:81 output.argmax(dim=1)
        assert np.array_equal(p, ref_p)
        preds.append(ref_p)
        targs.append(t)
    cm = conf.compute().numpy()
    assert np.array_equal(cm, M.confusion_matrix(np.concatenate(preds), np.concatenate(targs)))
    bad = EV.DeviceConfusion(dev)
    bad.update(torch.zeros(2, 5, device=dev), torch.tensor([1, 7]))
    with pytest.raises(ValueError):
        bad.compute()


@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["squeeze-ernet", "squeeze-redconv"])
def test_evaluate_model_matches_oracle(arch, tmp_path, dev):
    csv_path, rows, frames = _write_images(str(tmp_path), n=37, seed=8)
    sd = fixtures.get_state_dict(arch, "w3")
    model = rtdm_b200.from_state_dict(arch, sd, dev, "fp32")
    samples = EV.read_split(csv_path, str(tmp_path))
    metrics = EV.evaluate_model(model, EV.frame_batches(samples, batch_size=8, num_workers=4), dev)
    # oracle: PIL decode -> integer restatement of the eval transform -> fp64 forward -> argmax -> metrics
    x = np.concatenate([I.ingest(fr[None]) for fr in frames], 0)
    ref = E.forward(sd, x, arch, dtype=np.float64)
    pred = ref["probs"].argmax(1)
    targ = np.array([c for _, c in rows])
    srt = np.sort(ref["logits"], axis=1)
    assert ((srt[:, -1] - srt[:, -2]) > 1e-3 * np.abs(ref["logits"]).max()).all(), "test frames must have a clear top-1"
    cm = M.confusion_matrix(pred, targ)
    assert np.array_equal(metrics["confusion_matrix"], cm)
    want = {**M.micro_metrics(cm), **M.per_class_metrics(cm)}
    for k, v in want.items():
        assert metrics[k] == pytest.approx(v), k
    assert metrics["fps"] > 0 and metrics["avg_inference_time"] > 0
    # the reference's DataLoader form: float (B,3,140,140) tensors through model(data)
    loader = [(torch.from_numpy(x[i:i + 16]), torch.from_numpy(targ[i:i + 16])) for i in range(0, len(x), 16)]
    m2 = EV.evaluate_model(model, loader, dev)
    assert np.array_equal(m2["confusion_matrix"], cm)
    with pytest.raises(ValueError):
        EV.evaluate_model(model, [], dev)


# ------------------------------------------------------------------------------------ GPU: engine build / export
@pytest.mark.gpu
@pytest.mark.parametrize("arch,quant", [("squeeze-ernet", "fp32"), ("squeeze-ernet", "fp16"), ("squeeze-ernet", "int8"),
                                        ("squeeze-redconv", "fp16"), ("squeeze-redconv", "bf16"), ("squeeze-redconv", "int8"),
                                        ("ernet", "fp16")])
def test_build_trt_model_saves_and_reloads(arch, quant, tmp_path, dev):
    sd = fixtures.get_state_dict(arch, "shipped")
    src = rtdm_b200.from_state_dict(arch, sd, dev, "fp32")
    hw = BE.get_model_input_shape(arch)[2]
    args = argparse.Namespace(model=arch, quant=quant, output=None, output_dir=str(tmp_path))
    eng = BE.build_trt_model(src, torch.randn(1, 3, hw, hw), args)
    path = os.path.join(str(tmp_path), f"{arch}_{quant}_trt.pth")              # build_tensorrt_model.py:280-283
    assert eng.saved_path == path and os.path.exists(path)
    args2 = argparse.Namespace(model=arch, quant=quant, output="v2", output_dir=str(tmp_path))
    assert BE.build_trt_model(src, torch.randn(1, 3, hw, hw), args2).saved_path.endswith(f"{arch}_{quant}_v2_trt.pth")
    # the caller side, aider-predict.py:167-171
    trt_model = BE.TRTModule()
    trt_model.load_state_dict(torch.load(path, weights_only=True))
    trt_model = trt_model.to(dev)
    x = torch.from_numpy(fixtures.normal_tensors(6, seed=3, hw=hw)).to(dev)
    direct = rtdm_b200.from_state_dict(arch, sd, dev, quant)
    if quant == "int8":
        direct.set_act_scales(trt_model.act_scales)
    xin = x.half() if quant == "fp16" else x
    assert torch.equal(trt_model.logits(xin), direct.logits(xin))             # same blob -> same bits
    p = trt_model(xin)
    assert p.dtype == xin.dtype and tuple(p.shape) == (6, 5)
    # and against the oracle at the north_star tolerance of that precision
    if arch == "ernet":
        ref = E.forward_ernet(sd, x.cpu().numpy(), dtype=np.float64)["logits"]
    else:
        ref = E.forward(sd, x.cpu().numpy(), arch, dtype=np.float64)["logits"]
    got = trt_model.logits(xin).double().cpu().numpy()
    # int8: Gaussian-noise tensors lie outside the calibration distribution (frames); measured 0.11 (Squeeze_ErNET), 0.155 (RedConv)
    tol = {"fp32": 1e-4, "fp16": 2e-2, "bf16": 2e-2, "int8": 2e-1}[quant]
    assert np.abs(got - ref).max() / np.abs(ref).max() <= tol
    if quant != "int8":
        assert (got.argmax(1) == ref.argmax(1)).all()


# ------------------------------------------------------------------------------------ GPU: aider-predict.py mirror
@pytest.mark.gpu
@pytest.mark.parametrize("arch", ["squeeze-ernet", "ernet"])
def test_predict_matches_oracle(arch, tmp_path, dev):
    """predict(): class name and the reference's second-softmax confidence (aider-predict.py:76-84) from an image file."""
    from PIL import Image
    from rtdm_b200 import predict as P
    sd = fixtures.get_state_dict(arch, "w3")
    model = rtdm_b200.from_state_dict(arch, sd, dev, "fp32")
    size = 240 if arch == "ernet" else 140
    for i, (h, w) in enumerate([(240, 240), (300, 420)]):
        fr = fixtures.smooth_frames(1, h, w, seed=90 + i)[0]
        path = os.path.join(str(tmp_path), f"img{i}.png")
        Image.fromarray(fr).save(path)                                  # RGB on disk; cv2.imread returns it as BGR
        name, conf = P.predict(model, path, None, dev)
        x = I.ingest(fr[None], size)
        ref = (E.forward_ernet(sd, x, dtype=np.float64) if arch == "ernet" else E.forward(sd, x, arch, dtype=np.float64))["probs"][0]
        z = np.exp(ref - ref.max())
        second = z / z.sum()
        assert name == P.CLASSES[int(ref.argmax())]
        assert conf == pytest.approx(float(second[ref.argmax()]) * 100, rel=1e-4)
        # the reference's own sequence (PIL transform on the host, then model(data)) gives the same answer
        from oracle import ernet_torch as T
        from torchvision import transforms
        tf = transforms.Compose([transforms.Resize(int(size * 1.14)), transforms.CenterCrop(size), transforms.ToTensor(),
                                 transforms.Normalize(mean=[0.485, 0.456, 0.406], std=[0.229, 0.224, 0.225])])
        name2, conf2 = P.predict(model, path, tf, dev)
        assert name2 == name and conf2 == pytest.approx(conf, rel=1e-5)
    with pytest.raises(ValueError):
        P.predict(model, os.path.join(str(tmp_path), "missing.png"), None, dev)


@pytest.mark.gpu
def test_run_inference_stream_matches_batch_path(dev):
    """real-time-inference.py's per-frame call: FrameClassifier (pinned staging + graph replay) == forward_frames on the same
    BGR frames, frame after frame; the confidence is the reference's second softmax."""
    from rtdm_b200 import predict as P
    sd = fixtures.get_state_dict("squeeze-ernet", "w3")
    model = rtdm_b200.from_state_dict("squeeze-ernet", sd, dev, "bf16")
    frames = np.concatenate([fixtures.smooth_frames(5, 240, 240, seed=120), fixtures.noise_frames(2, seed=121)], 0)
    want = model.forward_frames(torch.from_numpy(frames).to(dev), bgr=True).cpu().numpy()
    clf = P.FrameClassifier(model, 240, 240)
    for i, fr in enumerate(frames):
        p = clf.probabilities(fr)
        assert np.array_equal(p, want[i])
        name, conf = P.run_inference(model, fr)
        z = np.exp(want[i] - want[i].max())
        assert name == P.CLASSES[int(want[i].argmax())]
        assert conf == pytest.approx(float(z[want[i].argmax()] / z.sum()) * 100, rel=1e-5)
    with pytest.raises(ValueError):
        clf(np.zeros((100, 100, 3), np.uint8))


@pytest.mark.gpu
def test_device_jpeg_decode_loader_matches_host_loader(tmp_path, dev):
    """frame_batches_device: files read on host threads, JPEGs decoded by one batched nvJPEG call per chunk, PNGs (not
    JPEG) through the host decoder.  Every sample exactly once, frames within a few grey levels of PIL's decode, and the
    evaluation result (confusion matrix) equal to the host-decode loop's up to near-tie predictions."""
    from PIL import Image
    rs = np.random.RandomState(5)
    frames = fixtures.smooth_frames(24, 240, 240, seed=9)
    samples = []
    for k in range(60):
        ext = "png" if k % 10 == 3 else "jpg"
        path = str(tmp_path / f"f{k:03d}.{ext}")
        img = frames[k % 24] if k % 7 else fixtures.smooth_frames(1, 180, 320, seed=k)[0]      # a second frame size
        Image.fromarray(img).save(path, quality=92) if ext == "jpg" else Image.fromarray(img).save(path)
        samples.append((path, int(rs.randint(0, 5))))
    host = {}
    for fb, tb in EV.frame_batches(samples, 16, 0, pin_memory=False):
        host.setdefault(tuple(fb.shape[1:3]), []).append((fb, tb))
    n = 0
    seen = {}
    for fb, tb in EV.frame_batches_device(samples, 16, dev, 2):
        assert fb.is_cuda and fb.dtype == torch.uint8 and fb.dim() == 4 and fb.shape[3] == 3 and fb.is_contiguous()
        n += fb.shape[0]
        seen.setdefault(tuple(fb.shape[1:3]), []).append((fb.cpu(), tb))
    assert n == len(samples) and set(seen) == set(host)
    for hw in host:
        a = torch.cat([f for f, _ in host[hw]]).to(torch.int16)
        b = torch.cat([f for f, _ in seen[hw]]).to(torch.int16)
        assert a.shape == b.shape
        assert torch.equal(torch.cat([t for _, t in host[hw]]), torch.cat([t for _, t in seen[hw]]))
        d = (a - b).abs()                                                            # nvJPEG vs libjpeg-turbo: IDCT rounding and chroma
        print(f"nvJPEG vs PIL {hw}: max {int(d.max())}, mean {float(d.float().mean()):.3f}, > 8 levels: {float((d > 8).float().mean()):.4f}")
        assert d.max() <= 48 and d.float().mean() < 3.0 and (d > 8).float().mean() < 0.05   # up-sampling differ (4:2:0 files)
    m = rtdm_b200.from_state_dict("squeeze-ernet", fixtures.get_state_dict("squeeze-ernet", "w3"), dev, "bf16")
    cm_h = EV.evaluate_model(m, EV.frame_batches(samples, 16, 2), dev)["confusion_matrix"]
    cm_d = EV.evaluate_model(m, EV.frame_batches_device(samples, 16, dev, 2), dev)["confusion_matrix"]
    assert int(np.abs(np.asarray(cm_h) - np.asarray(cm_d)).sum()) <= 2               # at most one prediction moved
