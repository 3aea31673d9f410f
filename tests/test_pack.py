"""Packer: blob structure, id agreement with the C header, and the algebraic folds (BN affine,
conv_red1 merge, head collapse) validated against the oracle on CPU."""
import os
import re

import numpy as np
import pytest

import fixtures
import packed_eval
import rtdm_b200.pack as P
from oracle import ernet_numpy as E

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_ids_match_header():
    hdr = open(os.path.join(ROOT, "real-time-disaster-management_b200", "csrc", "blob_format.h")).read()
    defs = {m.group(1): int(m.group(2), 0) for m in re.finditer(r"#define\s+ERNET_(\w+)\s+(0x[0-9A-Fa-f]+|\d+)u?", hdr)}
    assert defs["BLOB_MAGIC"] == P.MAGIC and defs["BLOB_VERSION"] == P.VERSION
    for name in ("T_STEM_W", "T_STEM_B", "T_BLOCK_BASE", "T_DW_W", "T_DW_B", "T_PW_W", "T_PW_B", "T_BN_S", "T_BN_T",
                 "T_RED2_W", "T_RED2_B", "T_RED3_W", "T_RED3_B", "T_HEAD_W", "T_HEAD_B", "T_TC_BASE"):
        assert defs[name] == getattr(P, name), name


@pytest.mark.parametrize("arch", fixtures.ARCHS)
@pytest.mark.parametrize("wset", ("shipped", "w3neg"))
def test_packed_form_matches_oracle(arch, wset):
    sd = fixtures.get_state_dict(arch, wset)
    blob = P.pack_state_dict(sd, arch, "fp32")
    x = fixtures.normal_tensors(2, seed=3)
    logits, taps = packed_eval.forward_packed(blob, x, arch)
    ref = E.forward(sd, x, arch, dtype=np.float64, want_taps=True)
    scale = np.abs(ref["logits"]).max()
    assert np.abs(logits - ref["logits"]).max() <= 2e-6 * scale        # weights were rounded to fp32 once
    for mine, theirs in (("stem", "stem"), ("pool1", "pool1"), ("pool2", "pool2"), ("pool3", "pool3")):
        r = np.transpose(ref["taps"][theirs], (0, 2, 3, 1))
        assert taps[mine].shape == r.shape
        assert np.abs(taps[mine] - r).max() <= 2e-6 * np.abs(r).max(), mine


def test_rejects_bad_state_dicts():
    sd = fixtures.get_state_dict("squeeze-ernet", "w3")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "ernet", "fp32")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "squeeze-redconv", "fp32")       # wrong key set
    bad = dict(sd)
    bad["conv1.weight"] = bad["conv1.weight"][:8]
    with pytest.raises(ValueError):
        P.pack_state_dict(bad, "squeeze-ernet", "fp32")
    with pytest.raises(ValueError):
        P.pack_state_dict(sd, "squeeze-ernet", "fp64")


def test_blob_layout():
    sd = fixtures.get_state_dict("squeeze-redconv", "w3")
    pz = packed_eval.parse_blob(P.pack_state_dict(sd, "squeeze-redconv", "fp32"))
    assert pz["arch"] == 1 and pz["precision"] == 0
    assert P.T_RED2_W in pz["tensors"] and P.T_HEAD_W in pz["tensors"]
    assert len(pz["tensors"][P.T_STEM_W][1]) == 27 * 8 * 4
