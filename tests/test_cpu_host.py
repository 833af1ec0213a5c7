"""CPU suite (no GPU): host logic of the drop-in module and the C-ABI surface."""
import ctypes
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT, load_golden


def test_library_loads_and_exports_every_declared_symbol():
    from cleanumamba_b200 import _lib
    lib = _lib.load()
    header = open(os.path.join(ROOT, "include", "cleanumamba_b200.h")).read()
    declared = set(re.findall(r"\b(cum_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/cleanumamba_b200.h but not exported"
    assert declared == set(_lib.EXPORTS), "ctypes table and header disagree"
    assert lib.cum_abi_version() == int(re.search(r"#define CUM_ABI_VERSION (\d+)", header).group(1))


def test_struct_layouts_match_header_field_order():
    from cleanumamba_b200 import _lib
    header = open(os.path.join(ROOT, "include", "cleanumamba_b200.h")).read()
    for struct, cls in (("cum_gemm_desc", _lib.GemmDesc), ("cum_scan_desc", _lib.ScanDesc)):
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (struct, struct), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.sub(r"\[.*\]", "", part.strip().split()[-1].lstrip("*")))
        assert names == [f[0] for f in cls._fields_], struct


def test_no_cpu_fallback():
    from cleanumamba_b200.network import Net
    fx = load_golden("tiny_equalwidth_seed0")
    net = Net("CleanUMamba", json.loads(fx["config"])).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU fallback"):
        net(fx["noisy"].clone())


def test_product_package_never_imports_oracle():
    pkg = os.path.join(ROOT, "cleanumamba_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "cleanumamba_oracle" not in src and "ref_loader" not in src and "ref_shim" not in src, f


def test_full_size_constructor_matches_reference_init_sums():
    """Seed-0 constructor == the reference constructor (per-tensor sums recorded by oracle/make_golden.py)."""
    from cleanumamba_b200.network import Net
    sums = json.load(open(os.path.join(GOLDEN, "full_init_seed0_sums.json")))
    for tag, ref in sums.items():
        torch.manual_seed(0)
        net = Net("CleanUMamba", ref["config"])
        sd = net.state_dict()
        assert list(sd.keys()) == list(ref["tensors"].keys())
        assert sum(p.numel() for p in net.parameters()) == ref["n_params"]
        for k, (shape, s, a) in ref["tensors"].items():
            assert list(sd[k].shape) == shape, k
            assert abs(float(sd[k].double().sum()) - s) <= 1e-6 * max(1.0, abs(a)), k


@pytest.mark.parametrize("name", ["e8_pruned_500k", "e6_pruned_200k", "mini_mamba_442k"])
def test_shipped_checkpoints_load_unchanged(name):
    from cleanumamba_b200.network import Net
    fx = load_golden(name)
    net = Net("CleanUMamba", json.loads(fx["config"]))
    net.load_pruned_state_dict(fx["state_dict"])
    sd = net.state_dict()
    assert set(sd.keys()) == set(fx["state_dict"].keys())   # (the mini experiment checkpoint lists norm before mixer)
    for k, v in fx["state_dict"].items():
        assert sd[k].shape == v.shape and torch.equal(sd[k].to(v.dtype), v)
    mx = net.tsfm_Mamba_layers[0].mixer
    assert mx.d_inner == mx.A_log.shape[0] and mx.d_state == mx.A_log.shape[1]
    assert net.encoder[1][0].in_channels == fx["state_dict"]["encoder.1.0.weight"].shape[1]
    assert net.decoder[0][2].out_channels == fx["state_dict"]["decoder.0.2.weight"].shape[1]


def test_shape_helpers_and_api_surface():
    from cleanumamba_b200 import CleanUMamba
    m = CleanUMamba(channels_H=8, max_H=16, encoder_n_layers=8, tsfm_n_layers=1, tsfm_d_model=16, tsfm_d_inner=32, tsfm_n_head=2)
    assert m.valid_length(160000) == 160254 and m.frame_length == 766 and m.total_stride == 256
    assert m.pad_signal(torch.zeros(1, 1, 160000)).shape[-1] == 160254
    cache = m.allocate_inference_cache(2, 1)
    assert cache[0][0].shape == (2, 32, 4) and cache[0][1].shape == (2, 32, 8)
    for name in ("forward", "feed", "flush", "load_pruned_state_dict", "allocate_inference_cache_layer"):
        assert callable(getattr(m, name))
    with pytest.raises(NotImplementedError):
        CleanUMamba(LSTM=True)
    with pytest.raises(ValueError):
        m.feed(torch.zeros(2, 3, 4))


def test_struct_sizes_and_offsets_match_a_c_compiler(tmp_path):
    """Compile the public header with plain gcc (it must be a C header: no CUDA / C++ needed) and compare sizeof / offsetof of
    every descriptor with the ctypes mirror in cleanumamba_b200/_lib.py."""
    import ctypes as C
    import os
    import shutil
    import subprocess
    from cleanumamba_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pairs = {"cum_gemm_desc": _lib.GemmDesc, "cum_scan_desc": _lib.ScanDesc, "cum_wgrad_desc": _lib.WgradDesc,
             "cum_scan_bwd_desc": _lib.ScanBwdDesc, "cum_enc0_block_desc": _lib.Enc0BlockDesc,
             "cum_dec_last_block_desc": _lib.DecLastBlockDesc, "cum_shift_entry": _lib.ShiftEntry}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "cleanumamba_b200.h"', 'int main(void) {']
    for cname, cls in pairs.items():
        lines.append(f'  printf("{cname} %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname}.{fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ['  return 0;', '}']
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    got = dict(l.split() for l in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(got[cname]) == C.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got[f"{cname}.{fname}"]) == getattr(cls, fname).offset, f"{cname}.{fname}"


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs next to the product arm): one JSON line with the product arm's
    metric / unit / workload and the keys of the bench contract."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                        "--seconds", "1"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    for k in ("impl", "metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
              "vs_baseline", "dtype", "data", "config", "cpu_baseline", "e2e"):
        assert k in d, k
    assert d["impl"] == "reference" and d["unit"] == "audio-s/s" and d["higher_is_better"] is True and d["value"] > 0
    assert "workload" in d["config"] and "offline forward" in d["config"]["workload"]
    # the unmodified reference (staged in baseline/_ref, or the /root/reference mount) when present, else the oracle port
    assert d["cpu_baseline"]["kind"] in ("reference", "port") and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    staged = os.path.isfile(os.path.join(root, "baseline", "_ref", "src", "network", "CleanUMamba.py")) or os.path.isdir("/root/reference/src")
    assert d["cpu_baseline"]["kind"] == ("reference" if staged else "port")
    assert d["cpu_baseline"]["port"]["value"] > 0           # the port is timed beside it
    assert set(d["config"]) == {"workload", "math", "global_batch", "clip_seconds", "parallelism", "l2"}     # == the product arm's config
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
