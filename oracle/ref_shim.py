"""Import the UNMODIFIED reference (ManifoldRG/NEKO) in this container.

TEST INFRASTRUCTURE ONLY.  Nothing in the product path (``neko_b200/``) imports this module.
It exists so that (1) ``oracle/make_golden.py`` can run the real reference on seeded inputs and
write the fixtures under ``tests/golden/`` and (2) ``tests/test_oracle_vs_reference.py`` can
pin the restatement in ``oracle/gato_oracle.py`` against the reference whenever
``/root/reference`` is mounted (it is NOT mounted on the GPU box) and (3) ``bench.py --impl reference`` /
``cpu_baseline`` can time the reference's own CPU path on the GPU box from the git-ignored archive
(``oracle/_ref/gato_ref.zip``, unpacked into a scratch directory of the process) that ``oracle/build_ref.py`` builds.

The reference was written for transformers 4.30.2 / torch 2.0.1 (``env.yml:8-36``); this image
has transformers 5.x and lacks gymnasium, so a handful of in-memory shims are installed before
``import gato.policy.gato_policy``.  No reference file is modified
or copied.
"""
from __future__ import annotations

import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
VENDORED_ROOT = os.path.join(_HERE, "_ref", "gato_ref.zip")     # written by oracle/build_ref.py (git-ignored; travels to the GPU box)


def _has_reference(root) -> bool:
    """A directory with gato/policy/gato_policy.py, or a zip archive with that member."""
    if not root:
        return False
    if os.path.isdir(root):
        return os.path.isfile(os.path.join(root, "gato", "policy", "gato_policy.py"))
    if os.path.isfile(root) and root.endswith(".zip"):
        import zipfile
        try:
            with zipfile.ZipFile(root) as z:
                return "gato/policy/gato_policy.py" in z.namelist()
        except zipfile.BadZipFile:
            return False
    return False


def _find_root() -> str:
    """The mounted reference if there is one, else the archive oracle/build_ref.py placed under oracle/_ref/."""
    for c in (os.environ.get("NEKO_REFERENCE_ROOT"), "/root/reference", VENDORED_ROOT):
        if _has_reference(c):
            return c
    return os.environ.get("NEKO_REFERENCE_ROOT", "/root/reference")


REFERENCE_ROOT = _find_root()


def reference_available() -> bool:
    return _has_reference(REFERENCE_ROOT)


def reference_kind() -> str:
    """'mounted' (/root/reference), 'vendored' (oracle/_ref) or 'absent'."""
    if not reference_available():
        return "absent"
    return "vendored" if os.path.abspath(REFERENCE_ROOT) == os.path.abspath(VENDORED_ROOT) else "mounted"


class _FakeTextTokenizer:
    """Only ``vocab_size`` is read on the hot path (gato_policy.py:57-60)."""

    def __init__(self, vocab_size: int = 50257):
        self.vocab_size = vocab_size

    def encode(self, s):  # pragma: no cover - not on the hot path
        raise NotImplementedError("offline shim: no GPT-2 BPE tables in this container")

    def decode(self, ids):  # pragma: no cover
        raise NotImplementedError("offline shim: no GPT-2 BPE tables in this container")


_installed = False
_text_vocab = 50257


def set_text_vocab(n: int) -> None:
    """Vocabulary size reported by the shimmed AutoTokenizer (50257 = GPT-2)."""
    global _text_vocab
    _text_vocab = int(n)


def install_shims() -> None:
    global _installed
    if _installed:
        return
    import transformers
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu

    # (1)(2) names trajectory_gpt2.py:37-43 imports from modeling_utils
    if not hasattr(mu, "Conv1D"):
        mu.Conv1D = pu.Conv1D
    for name in ("find_pruneable_heads_and_indices", "prune_conv1d_layer"):
        if not hasattr(mu, name):
            setattr(mu, name, getattr(pu, name, lambda *a, **k: None))
    if not hasattr(mu, "SequenceSummary"):
        import torch.nn as nn

        class SequenceSummary(nn.Module):  # dead code on the hot path
            def __init__(self, *a, **k):
                super().__init__()

        mu.SequenceSummary = SequenceSummary

    # (3) transformers.utils.model_parallel_utils (removed in HF 5)
    modname = "transformers.utils.model_parallel_utils"
    if modname not in sys.modules:
        try:
            __import__(modname)
        except Exception:
            m = types.ModuleType(modname)
            m.assert_device_map = lambda *a, **k: None
            m.get_device_map = lambda *a, **k: {}
            sys.modules[modname] = m

    # (4) gymnasium: only touched inside predict_control (gato_policy.py:564-567)
    if "gymnasium" not in sys.modules:
        try:
            import gymnasium  # noqa: F401
        except Exception:
            g = types.ModuleType("gymnasium")
            sp = types.ModuleType("gymnasium.spaces")

            class Box:  # noqa: D401
                pass

            class Discrete:
                pass

            class Env:
                pass

            sp.Box, sp.Discrete = Box, Discrete
            g.spaces, g.Env = sp, Env
            sys.modules["gymnasium"] = g
            sys.modules["gymnasium.spaces"] = sp

    # (5) AutoTokenizer.from_pretrained('gpt2') needs the hub; the hot path reads .vocab_size only
    class _AutoTok:
        @staticmethod
        def from_pretrained(name, *a, **k):
            return _FakeTextTokenizer(_text_vocab)

    transformers.AutoTokenizer = _AutoTok

    root = REFERENCE_ROOT
    if os.path.isfile(root) and root.endswith(".zip"):
        # the vendored archive: unpack into a scratch directory of this process (transformers 5 opens the source file of a model
        # class, which zipimport cannot serve); nothing is written into the repository tree
        import atexit
        import shutil
        import tempfile
        import zipfile
        scratch = tempfile.mkdtemp(prefix="neko_ref_")
        atexit.register(shutil.rmtree, scratch, ignore_errors=True)
        with zipfile.ZipFile(root) as z:
            z.extractall(scratch)
        root = scratch
    if root not in sys.path:
        sys.path.insert(0, root)

    import gato.transformers.trajectory_gpt2 as tg

    # (6) HF-4.30.2 behaviour of init_weights / get_head_mask for a model without tied weights
    tg.GPT2PreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
    tg.GPT2Model.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n

    # (7) --pretrained_lm (gato_policy.py:79-95): HF 5's from_pretrained needs post_init bookkeeping the reference's class does
    # not have, and with shim (6) its init pass re-initialises the weights it has just loaded.  Restore the HF-4.30.2
    # outcome: construct through HF, then put the checkpoint's tensors back (wpe / attn.bias are not part of the class).
    tg.GPT2PreTrainedModel.all_tied_weights_keys = {}
    _orig_from_pretrained = tg.GPT2Model.from_pretrained.__func__

    def _from_pretrained(cls, path, *a, **k):
        m = _orig_from_pretrained(cls, path, *a, **k)
        sd = load_gpt2_checkpoint(path)
        own = m.state_dict()
        m.load_state_dict({n: t for n, t in sd.items() if n in own and not n.endswith((".attn.bias", ".attn.masked_bias"))}, strict=False)
        # HF 5 builds the module on the meta device: the causal-mask buffers Attention.__init__ computes
        # (trajectory_gpt2.py:127-130) come back uninitialised.  Recompute them as __init__ does.
        import torch
        for blk in m.h:
            n_ctx = blk.attn.bias.shape[-1]
            blk.attn.bias = torch.tril(torch.ones((n_ctx, n_ctx), dtype=torch.uint8)).view(1, 1, n_ctx, n_ctx)
            blk.attn.masked_bias = torch.tensor(-1e4)
        return m

    tg.GPT2Model.from_pretrained = classmethod(_from_pretrained)
    _installed = True


def load_gpt2_checkpoint(path: str):
    """Tensors of a local HF GPT-2 checkpoint directory (model.safetensors or pytorch_model.bin), 'transformer.' prefix stripped."""
    import torch
    st = os.path.join(path, "model.safetensors")
    if os.path.exists(st):
        from safetensors.torch import load_file
        sd = load_file(st)
    else:
        sd = torch.load(os.path.join(path, "pytorch_model.bin"), map_location="cpu", weights_only=True)
    return {(n[len("transformer."):] if n.startswith("transformer.") else n): t for n, t in sd.items()}


def load_reference_policy_class():
    """Returns the reference's ``GatoPolicy`` class (gato/policy/gato_policy.py:18)."""
    if not reference_available():
        raise RuntimeError(f"reference tree not found under {REFERENCE_ROOT}")
    install_shims()
    from gato.policy.gato_policy import GatoPolicy

    return GatoPolicy
