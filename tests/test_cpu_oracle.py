"""CPU suite (no GPU): the oracle against the committed golden vectors and -- when the reference tree is mounted
(build container only) -- against the unmodified reference itself; plus the Mamba mixer against the independent
HuggingFace implementation."""
import json
import os

import pytest
import torch

import cleanumamba_oracle as orc
import ref_loader
from conftest import GOLDEN, load_golden

CKPT_FIXTURES = ["e8_pruned_500k", "e6_pruned_200k", "mini_mamba_442k"]


@pytest.mark.parametrize("name", CKPT_FIXTURES + ["tiny_equalwidth_seed0"])
def test_oracle_reproduces_reference_golden(name):
    fx = load_golden(name)
    y = orc.forward(fx["state_dict"], fx["noisy"])
    assert y.shape == fx["denoised"].shape
    assert (y - fx["denoised"]).abs().max().item() < 2e-6     # fp32, same ops in a different order


def test_oracle_fp64_brackets_fp32():
    fx = load_golden("mini_mamba_442k")
    y64 = orc.forward(fx["state_dict"], fx["noisy"], dtype=torch.float64)
    assert (y64.float() - fx["denoised"]).abs().max().item() < 5e-6


def test_oracle_does_not_mutate_input():
    fx = load_golden("tiny_equalwidth_seed0")
    x = fx["noisy"].clone()
    orc.forward(fx["state_dict"], x)
    assert torch.equal(x, fx["noisy"])


def test_stream_oracle_pins_to_reference_feed_in_bug_compat_mode():
    """The shipped feed() (skip-order bug at CleanUMamba.py:474 included) is reproduced bit-for-bit on the only kind
    of model it can run on (equal widths)."""
    fx = load_golden("tiny_equalwidth_seed0")
    so = orc.StreamOracle(fx["state_dict"], compat_skip_order_bug=True)
    a = so.feed(fx["noisy"][:, 0])
    b = so.flush()
    assert torch.equal(a, fx["stream_feed"]) and torch.equal(b, fx["stream_flush"])


@pytest.mark.parametrize("name", ["tiny_equalwidth_seed0", "e6_pruned_200k"])
def test_stream_oracle_equals_offline_forward(name):
    """With the intended skip order, every emitted sample equals offline forward (normalize_input=False)."""
    fx = load_golden(name)
    x = fx["noisy"][:1, 0, :3000]
    so = orc.StreamOracle(fx["state_dict"], normalize_input=False)
    outs = [so.feed(x[:, i:i + 700]) for i in range(0, x.shape[1], 700)]
    s = torch.cat(outs, 1)
    off = orc.forward(fx["state_dict"], x, normalize_input=False)[:, 0]
    assert s.shape[1] > 0 and s.shape[1] % so.total_stride == 0
    assert (s - off[:, : s.shape[1]]).abs().max().item() < 2e-6


def test_stream_oracle_batched_streams_are_independent():
    fx = load_golden("tiny_equalwidth_seed0")
    x = torch.randn(3, 1200, generator=torch.Generator().manual_seed(0)) * 0.1
    sb = orc.StreamOracle(fx["state_dict"], batch=3).feed(x)
    for i in range(3):
        si = orc.StreamOracle(fx["state_dict"], batch=1).feed(x[i:i + 1])
        assert (sb[i:i + 1] - si).abs().max().item() < 1e-6


def test_valid_length_known_answers():
    # 160000 -> 160254 (E8) / 160062 (E6); frame length 766 / 190
    assert orc.valid_length(160000, 8) == 160254 and orc.valid_length(160000, 6) == 160062
    assert orc.valid_length(1, 8) == 766 and orc.valid_length(1, 6) == 190


def test_selective_scan_against_definition():
    """Brute-force per-element recurrence (pure Python loops, tiny case)."""
    g = torch.Generator().manual_seed(1)
    b, d, l, n = 1, 3, 6, 4
    u, delta = torch.randn(b, d, l, generator=g), torch.randn(b, d, l, generator=g)
    A = -torch.rand(d, n, generator=g)
    Bm, Cm = torch.randn(b, n, l, generator=g), torch.randn(b, n, l, generator=g)
    D, z, bias = torch.randn(d, generator=g), torch.randn(b, d, l, generator=g), torch.randn(d, generator=g)
    y = orc.selective_scan(u.double(), delta.double(), A.double(), Bm.double(), Cm.double(), D.double(), z.double(),
                           bias.double(), True)
    import math
    for di in range(d):
        h = [0.0] * n
        for t in range(l):
            dl = math.log1p(math.exp(float(delta[0, di, t]) + float(bias[di])))
            acc = 0.0
            for k in range(n):
                h[k] = math.exp(dl * float(A[di, k])) * h[k] + dl * float(Bm[0, k, t]) * float(u[0, di, t])
                acc += h[k] * float(Cm[0, k, t])
            acc += float(D[di]) * float(u[0, di, t])
            zz = float(z[0, di, t])
            acc *= zz / (1 + math.exp(-zz))
            assert abs(acc - float(y[0, di, t])) < 1e-9


def test_mixer_matches_huggingface_slow_forward():
    """Independent cross-check of the restated dependency: transformers' MambaMixer.slow_forward."""
    tr = pytest.importorskip("transformers")
    from transformers.models.mamba.configuration_mamba import MambaConfig
    from transformers.models.mamba.modeling_mamba import MambaMixer
    cfg = MambaConfig(hidden_size=64, state_size=16, conv_kernel=4, expand=2, time_step_rank=4, use_bias=False,
                      use_conv_bias=True, num_hidden_layers=1)
    torch.manual_seed(0)
    hf = MambaMixer(cfg, layer_idx=0).eval()
    sd = {"m." + k: v for k, v in hf.state_dict().items()}
    x = torch.randn(2, 50, 64)
    with torch.no_grad():
        ref = hf.slow_forward(x)
    got = orc.mamba_mixer(x, sd, "m.")
    assert (got - ref).abs().max().item() < 1e-5
    del tr


needs_ref = pytest.mark.skipif(not ref_loader.reference_available(), reason="reference tree not mounted (GPU box)")


@needs_ref
@pytest.mark.parametrize("rel", ["checkpoints/pruned/CleanUMamba-3N-E8_pruned-200K.pkl",
                                 "checkpoints/pruned/CleanUMamba-3N-E6_pruned-500k.pkl"])
def test_oracle_vs_live_reference_on_other_checkpoints(rel):
    """Checkpoints that are NOT among the committed fixtures, run through the unmodified reference here."""
    refnet = ref_loader.import_reference()
    ck = torch.load(os.path.join(ref_loader.REFERENCE_ROOT, rel), map_location="cpu", weights_only=True)
    m = refnet.Net("CleanUMamba", dict(ck["network_config"]))
    m.load_pruned_state_dict(ck["model_state_dict"])
    m.float().eval()
    _, noisy = orc.synth_batch(1, 0.5, seed=99)
    with torch.no_grad():
        ref = m(noisy.clone())
    got = orc.forward(ck["model_state_dict"], noisy)
    assert (got - ref).abs().max().item() < 2e-6


@needs_ref
def test_product_constructor_reproduces_reference_init():
    """Same seed -> identical weights as the reference constructor (E6 config; E8 is covered by the sums file)."""
    from cleanumamba_b200.network import Net
    refnet = ref_loader.import_reference()
    cfg = json.load(open(os.path.join(ref_loader.REFERENCE_ROOT, "configs/exp/models/DNS-CleanUMamba-3N-E6.json")))
    torch.manual_seed(0)
    a = refnet.Net(cfg["network"], cfg["network_config"]).state_dict()
    torch.manual_seed(0)
    b = Net(cfg["network"], cfg["network_config"]).state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert torch.equal(a[k], b[k]), k


@needs_ref
def test_loss_restatement_matches_live_reference():
    """cleanumamba_b200.loss vs the reference's own MultiResolutionSTFTLoss (src/util/stft_loss.py) on the same input."""
    import sys
    ref_loader.import_reference()
    from src.util.stft_loss import MultiResolutionSTFTLoss as RefLoss
    from cleanumamba_b200.loss import DEFAULT_STFT_CONFIG, MultiResolutionSTFTLoss
    g = torch.Generator().manual_seed(3)
    x, y = torch.randn(2, 1, 8000, generator=g) * 0.1, torch.randn(2, 1, 8000, generator=g) * 0.1
    for band in ("full", "high"):
        cfg = dict(DEFAULT_STFT_CONFIG, band=band)
        a = MultiResolutionSTFTLoss(**cfg)(x, y)
        b = RefLoss(**cfg)(x, y)
        assert torch.allclose(a[0], b[0], rtol=1e-5) and torch.allclose(a[1], b[1], rtol=1e-5), band
    del sys


def test_channel_importances_restatement_matches_live_reference():
    """oracle.channel_importances == the reference's PruningModule.channel_importances (pruninggroup.py:160-226) on conv / linear /
    vector parameters, both dims, with a channel offset and several heads.  (src.pruning.util needs absent packages: only the
    one symbol pruninggroup imports from it is stubbed.)"""
    import sys
    import types
    import ref_loader
    if not os.path.isdir(os.path.join("/root/reference", "src", "pruning")):
        pytest.skip("reference tree not mounted")
    ref_loader.import_reference()
    if "src.pruning.util" not in sys.modules:
        stub = types.ModuleType("src.pruning.util")
        stub.prune_parameter_and_grad = lambda *a, **k: None
        sys.modules["src.pruning.util"] = stub
    from src.pruning.pruninggroup import PruningModule
    g = torch.Generator().manual_seed(0)
    cases = [(torch.nn.Conv1d(12, 16, 4), 0, 0, 16, 1), (torch.nn.Conv1d(12, 16, 4), 1, 0, 12, 1), (torch.nn.Linear(10, 24), 0, 8, 4, 4),
             (torch.nn.Linear(10, 24), 1, 2, 8, 1), (torch.nn.LayerNorm(14), 0, 0, 14, 1)]
    for mod, dim, off, n_ch, heads in cases:
        mod.weight.grad = torch.randn(mod.weight.shape, generator=g)
        pm = PruningModule(mod, dim=dim, n_heads=heads, channel_offset=off, statistics=False)
        pm.group = types.SimpleNamespace(n_channels=n_ch)
        if off or n_ch * heads != mod.weight.shape[dim]:        # the rest of the matrix belongs to a following module
            pm.next_module_to_offset = types.SimpleNamespace(channel_offset=off + n_ch * heads, module=mod)
        want = pm.channel_importances()
        got = orc.channel_importances(mod.weight, mod.weight.grad, dim, off, n_ch, heads)
        for k in ("weight", "grad", "taylor_individual", "taylor_squared_individual", "taylor_group"):
            assert torch.allclose(got[k], want[k], rtol=1e-5, atol=1e-7), (type(mod).__name__, dim, k)
        assert got["n_parameters"] == want["n_parameters"]
