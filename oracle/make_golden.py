"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.pt by running the UNMODIFIED reference
(/root/reference, imported through oracle/ref_loader.py on top of oracle/ref_shim) in the build container.

    python oracle/make_golden.py          # rewrites tests/golden/

Each fixture is a dict of plain tensors (loadable with ``weights_only=True``):
  state_dict   the checkpoint's ``model_state_dict`` exactly as shipped (fp16 tensors, reference key names)
  config       ``network_config`` of the checkpoint (JSON string)
  noisy        (B,1,T) fp32 synthetic input (oracle.synth_batch, seed 1234)
  clean        (B,1,T) fp32 clean target of the mixture (for delta-SI-SDR)
  denoised     (B,1,T) fp32 output of reference ``CleanUMamba.forward`` after ``model.float()``
                (loading recipe of src/examples/loading_pretrained_models.py:7-19)
  stream_*     outputs of reference ``feed`` / ``flush`` where the shipped code can run (equal-width model only:
                the skip-order bug at CleanUMamba.py:474 crashes on every shipped checkpoint)
Checkpoints travel as fixtures (they are data, not source); nothing else is copied from the reference.
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import cleanumamba_oracle as orc  # noqa: E402
import ref_loader  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
CKPTS = {
    "e8_pruned_500k": "checkpoints/pruned/CleanUMamba-3N-E8_pruned-500K.pkl",      # BASELINE config 1 model
    "e6_pruned_200k": "checkpoints/pruned/CleanUMamba-3N-E6_pruned-200k.pkl",
    "mini_mamba_442k": "checkpoints/experiments/Experiment_CleanU_Mamba.pkl",
}


def main():
    refnet = ref_loader.import_reference()
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count())
    for name, rel in CKPTS.items():
        ck = torch.load(os.path.join(ref_loader.REFERENCE_ROOT, rel), map_location="cpu", weights_only=True)
        model = refnet.Net("CleanUMamba", dict(ck["network_config"]))
        model.load_pruned_state_dict(ck["model_state_dict"])
        model.float().eval()
        clean, noisy = orc.synth_batch(2, 0.75)
        with torch.no_grad():
            out = model(noisy.clone())
        fx = dict(state_dict={k: v.clone() for k, v in ck["model_state_dict"].items()},
                  config=json.dumps(ck["network_config"]), noisy=noisy, clean=clean, denoised=out)
        torch.save(fx, os.path.join(OUT, f"{name}.pt"))
        print(name, "params", sum(p.numel() for p in model.parameters()), "out rms", out.pow(2).mean().sqrt().item())

    # random-init regular model (reference constructor, seed 0): covers init parity + the streaming path
    cfg = dict(channels_H=16, max_H=16, encoder_n_layers=4, tsfm_n_layers=2, tsfm_n_head=2, tsfm_d_model=16,
               tsfm_d_inner=32)
    torch.manual_seed(0)
    model = refnet.Net("CleanUMamba", dict(cfg)).eval()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    _, noisy = orc.synth_batch(1, 0.2)
    with torch.no_grad():
        out = model(noisy.clone())
        s1 = model.feed(noisy[:, 0].clone())
        s2 = model.flush()
    torch.save(dict(state_dict=sd, config=json.dumps(cfg), noisy=noisy, denoised=out, stream_feed=s1,
                    stream_flush=s2), os.path.join(OUT, "tiny_equalwidth_seed0.pt"))

    # constructor / init parity: per-tensor float64 sums of the seeded full-size models (weights too big to ship)
    sums = {}
    for tag in ("DNS-CleanUMamba-3N-E8", "DNS-CleanUMamba-3N-E6"):
        cfgj = json.load(open(os.path.join(ref_loader.REFERENCE_ROOT, "configs/exp/models", tag + ".json")))
        torch.manual_seed(0)
        m = refnet.Net(cfgj["network"], cfgj["network_config"])
        sums[tag] = {"config": cfgj["network_config"], "n_params": sum(p.numel() for p in m.parameters()),
                     "tensors": {k: [list(v.shape), float(v.double().sum()), float(v.double().abs().sum())]
                                 for k, v in m.state_dict().items()}}
    json.dump(sums, open(os.path.join(OUT, "full_init_seed0_sums.json"), "w"))
    print({k: v["n_params"] for k, v in sums.items()})


if __name__ == "__main__":
    main()
