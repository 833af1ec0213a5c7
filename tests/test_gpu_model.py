"""End-to-end GPU parity of CleanUMamba.forward against the committed golden vectors (reference outputs) and the
CPU oracle.  Bar (BASELINE.json north_star): max-abs <= 1e-4 on the waveform and |delta SI-SDR| <= 0.01 dB (fp32)."""
import json

import pytest
import torch

import cleanumamba_oracle as orc
from conftest import load_golden

pytestmark = pytest.mark.gpu

TOL_MAXABS = 1e-4
TOL_SISDR_DB = 0.01


def build(fx, **kw):
    from cleanumamba_b200.network import Net
    net = Net("CleanUMamba", {**json.loads(fx["config"]), **kw})
    if any(v.shape != net.state_dict()[k].shape for k, v in fx["state_dict"].items()):
        net.load_pruned_state_dict(fx["state_dict"])
    else:
        net.load_state_dict(fx["state_dict"])
    return net.cuda().float().eval()


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("name", ["e8_pruned_500k", "e6_pruned_200k", "mini_mamba_442k", "tiny_equalwidth_seed0"])
def test_forward_matches_reference_golden(name, math):
    fx = load_golden(name)
    net = build(fx, math_mode=math)
    x = fx["noisy"].cuda()
    with torch.no_grad():
        y = net(x)
    ref = fx["denoised"]
    assert y.shape == ref.shape
    err = (y.cpu() - ref).abs().max().item()
    assert err <= TOL_MAXABS, f"max-abs {err}"
    assert err / ref.pow(2).mean().sqrt().item() < 1e-4          # relative to the output rms (SURVEY §8d caveat)
    if "clean" in fx:
        d = (orc.si_sdr(y.cpu(), fx["clean"]) - orc.si_sdr(ref, fx["clean"])).abs().max().item()
        assert d <= TOL_SISDR_DB
    # the reference normalises the caller's tensor in place (CleanUMamba.py:262): so do we
    std = fx["noisy"].std(dim=2, keepdim=True) + 1e-3
    assert torch.allclose(x.cpu(), fx["noisy"] / std, rtol=1e-5, atol=1e-7)


def test_forward_2d_input_and_skip_outputs():
    fx = load_golden("mini_mamba_442k")
    net = build(fx)
    x = fx["noisy"][:, 0].cuda()
    with torch.no_grad():
        y, skips = net(x.clone(), return_skip_connections=True)
    ref, inter = orc.forward(fx["state_dict"], fx["noisy"], return_intermediates=True)
    assert (y.cpu() - ref).abs().max().item() <= TOL_MAXABS
    assert len(skips) == len(inter["skips"]) + 1
    for got, want in zip(skips[:-1], inter["skips"][::-1]):
        assert got.shape == want.shape and (got.cpu() - want).abs().max().item() < 1e-4


def test_normalize_input_false_returns_padded_length():
    fx = load_golden("tiny_equalwidth_seed0")
    net = build(fx, normalize_input=False)
    x = fx["noisy"].cuda()
    with torch.no_grad():
        y = net(x)
    ref = orc.forward(fx["state_dict"], fx["noisy"], normalize_input=False)
    assert y.shape == ref.shape and y.shape[-1] == net.valid_length(x.shape[-1])
    assert (y.cpu() - ref).abs().max().item() <= TOL_MAXABS
    assert torch.equal(x.cpu(), fx["noisy"])


@pytest.mark.parametrize("math", ["fp32", "tf32x3"])
@pytest.mark.parametrize("cfg_name,seconds", [("DNS-CleanUMamba-3N-E8", 1.0), ("DNS-CleanUMamba-3N-E6", 0.5)])
def test_full_size_random_init_matches_oracle(cfg_name, seconds, math):
    """E8-full / E6-high (random init, seed 0 == reference constructor; checkpoints are not shipped)."""
    from cleanumamba_b200.network import Net
    sums = json.load(open(__import__("os").path.join(__import__("conftest").GOLDEN, "full_init_seed0_sums.json")))[cfg_name]
    torch.manual_seed(0)
    net = Net("CleanUMamba", dict(sums["config"], math_mode=math))
    sd = {k: v.clone() for k, v in net.state_dict().items()}
    net = net.cuda().eval()
    clean, noisy = orc.synth_batch(2, seconds)
    ref = orc.forward(sd, noisy)
    with torch.no_grad():
        y = net(noisy.cuda())
    err = (y.cpu() - ref).abs().max().item()
    rms = ref.pow(2).mean().sqrt().item()
    print(f"\n[{cfg_name} {math}] max-abs {err:.3e}  out-rms {rms:.3e}")
    assert err <= TOL_MAXABS, f"max-abs {err} (rms {rms})"
    d = (orc.si_sdr(y.cpu(), clean) - orc.si_sdr(ref, clean)).abs().max().item()
    assert d <= TOL_SISDR_DB
