"""Replays tests/golden/reference_known_answers.json (the reference's own known-answer vectors)
against the CPU oracle and -- on a GPU -- through the C ABI."""
import json
import os

import numpy as np
import pytest

import oracle as orc

CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["cases"]
IDS = [f"{i}:{c['filter']}" for i, c in enumerate(CASES)]


def _oracle(case):
    x = np.asarray(case["input"], np.float32)
    if case["filter"] == "fir":
        return orc.fir(case["taps"], x, case["out_cap"])
    if case["filter"] == "decimating_fir":
        return orc.decim_fir(case["taps"], case["decim"], x, case["out_cap"])
    return orc.resamp_fir(case["taps"], case["interp"], case["decim"], x, case["out_cap"])


@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_oracle_reproduces_reference_vector(case):
    c, p, st, o = _oracle(case)
    assert (c, p, int(st)) == (case["consumed"], case["produced"], case["status"]), case["cite"]
    assert list(map(float, o)) == list(map(float, case["output"])), case["cite"]


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES, ids=IDS)
def test_device_reproduces_reference_vector(case):
    import torch
    import futuresdr_b200 as fb
    taps = np.asarray(case["taps"], np.float32)
    if case["filter"] == "fir":
        f = fb.FirFilter(taps, sample_dtype=np.float32)
    elif case["filter"] == "decimating_fir":
        f = fb.DecimatingFirFilter(case["decim"], taps, sample_dtype=np.float32)
    else:
        f = fb.PolyphaseResamplingFir(case["interp"], case["decim"], taps, sample_dtype=np.float32)
    x = torch.from_numpy(np.asarray(case["input"], np.float32)).cuda()
    cap = case["out_cap"]
    out = torch.full((max(cap, 1),), 7.0, dtype=torch.float32, device="cuda")[:cap]
    c, p, st = f.filter(x, out)
    torch.cuda.synchronize()
    assert (c, p, int(st)) == (case["consumed"], case["produced"], case["status"]), case["cite"]
    assert list(map(float, out[:p].cpu().numpy())) == list(map(float, case["output"])), case["cite"]


# ---- block-level known answers (tests/moving_avg.rs): MovingAvg is PINNED by the reference's own vectors ------------
BLOCK_CASES = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "reference_known_answers.json")))["block_cases"]


def _f32(vals):
    return np.asarray([float(v) for v in vals], np.float32)       # "nan" / "inf" strings -> non-finite floats


@pytest.mark.parametrize("case", BLOCK_CASES, ids=[c["cite"].split()[-1] for c in BLOCK_CASES])
def test_oracle_reproduces_reference_block_vector(case):
    m = orc.MovingAvg(case["width"], case["decay_factor"], case["history_size"])
    c, p, o = m.work(_f32(case["input"]), case["out_cap"])
    assert p == len(case["output"])
    assert np.array_equal(o, np.asarray(case["output"], np.float32)), case["cite"]     # assert_eq! on Vec<f32>


@pytest.mark.gpu
@pytest.mark.parametrize("case", BLOCK_CASES, ids=[c["cite"].split()[-1] for c in BLOCK_CASES])
def test_device_reproduces_reference_block_vector(case):
    import torch
    from futuresdr_b200.blocks import Mocker, MovingAvg
    blk = MovingAvg(case["width"], case["decay_factor"], case["history_size"])
    m = Mocker(blk)
    m.input(_f32(case["input"]))
    m.init_output(case["out_cap"])
    m.run()
    torch.cuda.synchronize()
    got = m.output().cpu().numpy()
    assert np.array_equal(got, np.asarray(case["output"], np.float32)), case["cite"]
