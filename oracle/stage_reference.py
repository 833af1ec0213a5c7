"""TEST / BASELINE INFRASTRUCTURE ONLY -- stages the files of the UNMODIFIED reference that its CPU forward needs into the
git-ignored ``baseline/_ref/`` so that they travel to the GPU box (which has no /root/reference) together with the snapshot:

    src/__init__.py, src/network/{CleanUMamba,layers,network}.py, src/util/{util,stft_loss}.py

Byte-for-byte copies made at build time by ``__graft_entry__.build()`` (only where /root/reference is mounted), never
committed (``baseline/_ref/`` is in .gitignore), never imported by the product package.  ``bench.py --impl reference`` and the
``cpu_baseline`` leg import them through ``oracle/ref_loader.py`` on top of ``oracle/ref_shim`` (the stand-in for the absent
``mamba_ssm`` wheel: ``selective_scan_ref`` and the slow-path ``Mamba`` / ``Block`` of mamba-ssm 1.2.2)."""
import filecmp
import os
import shutil

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get("CLEANUMAMBA_REFERENCE_SRC", "/root/reference")
DST = os.path.join(ROOT, "baseline", "_ref")
FILES = ["src/__init__.py", "src/network/CleanUMamba.py", "src/network/layers.py", "src/network/network.py",
         "src/util/util.py", "src/util/stft_loss.py"]


def stage(verbose: bool = False) -> bool:
    """Copy the reference files when the mount is present; returns True when baseline/_ref is complete afterwards."""
    if os.path.isfile(os.path.join(SRC, FILES[1])):
        for rel in FILES:
            src, dst = os.path.join(SRC, rel), os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            if not (os.path.isfile(dst) and filecmp.cmp(src, dst, shallow=False)):
                shutil.copyfile(src, dst)
                if verbose:
                    print(f"staged {rel}")
    return all(os.path.isfile(os.path.join(DST, rel)) for rel in FILES)


if __name__ == "__main__":
    print("baseline/_ref complete:", stage(verbose=True))
