"""TEST / BENCH INFRASTRUCTURE ONLY -- recipe for oracle/_ref/.

The reference is pure Python (no native code to compile), so "building" it for the GPU box means placing an UNMODIFIED copy of
the source files of this path where bench.py's reference arms can import them: /root/reference exists only in the authoring
container, oracle/_ref/ is git-ignored (never part of the history) but not gpurun-ignored (it travels with the snapshot).

    python -m oracle.build_ref          # copies modules/{mage_model,vqvae_model}.py and utils/util.py, prints their sha256

Nothing under mage_b200/ may import oracle/_ref (tests/test_abi.py greps for it); it is loaded only through oracle/ref_shims.py.
"""
from __future__ import annotations

import hashlib
import os
import shutil

SRC = os.environ.get("MAGE_REFERENCE_SRC", "/root/reference")
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
FILES = ("modules/mage_model.py", "modules/vqvae_model.py", "utils/util.py")


def build(verbose: bool = True) -> bool:
    """Returns True if oracle/_ref/ is (now) populated."""
    if not os.path.isfile(os.path.join(SRC, FILES[0])):
        return os.path.isfile(os.path.join(DST, FILES[0]))
    for rel in FILES:
        dst = os.path.join(DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC, rel), dst)
        if verbose:
            print(f"oracle/_ref/{rel}  sha256 {hashlib.sha256(open(dst, 'rb').read()).hexdigest()[:16]}")
    return True


if __name__ == "__main__":
    build()
