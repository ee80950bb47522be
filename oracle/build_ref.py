"""Recipe for oracle/_ref/: the reference's own hot-path package, unmodified, where bench.py's CPU arm can import it.

TEST / MEASUREMENT INFRASTRUCTURE ONLY -- nothing under neko_b200/ imports it.

The GPU box has no /root/reference, and the reference cannot be pip-installed into a target directory: its setup.py
declares ``packages=['gato']`` only, so ``pip install --target ... /root/reference`` ships gato/__init__.py and none of
the sub-packages (tried: DESIGN.md section 1).  The reference is pure Python, so "building" it is placing the four files
of the path (gato/policy/{gato_policy,embeddings,input_tokenizers}.py, gato/transformers/trajectory_gpt2.py + the
package __init__ files) under oracle/_ref/, byte for byte.  oracle/_ref/ is git-ignored (never part of the history)
but not gpurun-ignored, so it travels to the box like the built .so; ``__graft_entry__.build()`` runs this whenever
/root/reference is present.  The in-memory shims of oracle/ref_shim.py are applied at import time as before."""
from __future__ import annotations

import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, "_ref")
FILES = ["gato/__init__.py", "gato/policy/__init__.py", "gato/policy/gato_policy.py", "gato/policy/embeddings.py",
         "gato/policy/input_tokenizers.py", "gato/transformers/__init__.py", "gato/transformers/trajectory_gpt2.py"]


def build(src: str = "/root/reference") -> bool:
    """Returns True when oracle/_ref holds the reference's path files (copied now or already there)."""
    if not os.path.isfile(os.path.join(src, "gato", "policy", "gato_policy.py")):
        return os.path.isfile(os.path.join(DST, "gato", "policy", "gato_policy.py"))
    manifest = []
    for rel in FILES:
        s, d = os.path.join(src, rel), os.path.join(DST, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        if os.path.exists(s):
            shutil.copyfile(s, d)
            manifest.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {rel}")
        else:                      # a package without an __init__.py in the reference
            open(d, "w").close()
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("\n".join(manifest) + "\n")
    return True


if __name__ == "__main__":
    ok = build(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("oracle/_ref:", "ready" if ok else "reference tree not found")
